// fp64_rates.cu -- pipe-rate probes that decide the likelihood kernel's arithmetic:
// DFMA, DMMA m8n8k4 (FP64 tensor core), F2F f32->f64, MUFU.RCP64H, integer-trick f32->f64 widening.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float *out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_cvt(double *out, const float *in) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + i];
    double s = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += (double)v[i];  // F2F + DADD
            v[i] += 1.0f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// f32 -> f64 for normal numbers with integer ops only (no XU pipe)
__device__ __forceinline__ double widen_bits(float f) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t hi = ((u >> 3) & 0x0FFFFFFFu) + 0x38000000u | (u & 0x80000000u);
    const uint32_t lo = u << 29;
    return __hiloint2double((int)hi, (int)lo);
}
__global__ void k_cvt_int(double *out, const float *in) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + i];
    double s = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += widen_bits(v[i]);
            v[i] += 1.0f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_rcp64h(double *out, double a) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v[i]));
            v[i] = r;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


__global__ void k_rcp_acc(double *out) {
    // max relative error of the rcp.approx.ftz.f64 seed and of one / two Newton steps, over a sweep of mantissas/exponents
    double worst0 = 0, worst1 = 0, worst2 = 0;
    for (int i = 0; i < 4096; ++i) {
        const double p = ldexp(1.0 + (threadIdx.x * 4096 + i) / (4096.0 * blockDim.x), (int)blockIdx.x - 80);
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
        const double ex = 1.0 / p;
        worst0 = fmax(worst0, fabs(r - ex) / ex);
        double e = fma(-p, r, 1.0); double r1 = fma(r, e, r);
        worst1 = fmax(worst1, fabs(r1 - ex) / ex);
        e = fma(-p, r1, 1.0); double r2 = fma(r1, e, r1);
        worst2 = fmax(worst2, fabs(r2 - ex) / ex);
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 0] = worst0;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 1] = worst1;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 2] = worst2;
}

__global__ void k_dmma_lat(double *out, double a, double b, long long *cyc) {
    double c0 = threadIdx.x, c1 = 1;
    long long t0 = clock64();
    for (int it = 0; it < 1024; ++it) dmma(c0, c1, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_lat(double *out, double a, double b, long long *cyc) {
    double c0 = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < 1024; ++it) c0 = fma(c0, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = c0;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; float *fin;
    const int TPB = 256;
    for (int cps : {1, 2, 4}) {
        const int blocks = sms * cps;
        cudaMalloc(&out, sizeof(double) * blocks * TPB);
        cudaMalloc(&fin, sizeof(float) * (TPB + 16));
        cudaMemset(fin, 0, sizeof(float) * (TPB + 16));
        const double thr = (double)blocks * TPB;
        float ms;
        ms = time_it([&] { k_dfma<<<blocks, TPB>>>(out, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DFMA      %8.3f ms  %7.2f TFLOP/s  (%.1f fma/clk/SM @1.965GHz)\n", cps, ms, thr * 16 * ITERS * 2 / ms / 1e9,
               thr * 16 * ITERS / (ms * 1e-3) / sms / 1.965e9);
        ms = time_it([&] { k_ffma<<<blocks, TPB>>>((float *)out, 1.0000001f, 1e-9f); });
        printf("ctas/sm %d  FFMA      %8.3f ms  %7.2f TFLOP/s\n", cps, ms, thr * 16 * ITERS * 2 / ms / 1e9);
        ms = time_it([&] { k_dmma<4><<<blocks, TPB>>>(out, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DMMA x4   %8.3f ms  %7.2f TFLOP/s  (%.1f fma/clk/SM)\n", cps, ms, thr / 32 * 4 * ITERS * 512.0 / ms / 1e9,
               thr / 32 * 4 * ITERS * 256.0 / (ms * 1e-3) / sms / 1.965e9);
        ms = time_it([&] { k_dmma<8><<<blocks, TPB>>>(out, 1.0000001, 1e-9); });
        printf("ctas/sm %d  DMMA x8   %8.3f ms  %7.2f TFLOP/s  (%.1f fma/clk/SM)\n", cps, ms, thr / 32 * 8 * ITERS * 512.0 / ms / 1e9,
               thr / 32 * 8 * ITERS * 256.0 / (ms * 1e-3) / sms / 1.965e9);
        ms = time_it([&] { k_cvt<<<blocks, TPB>>>(out, fin); });
        printf("ctas/sm %d  F2F+DADD  %8.3f ms  %7.2f Gcvt/s  (%.1f /clk/SM)\n", cps, ms, thr * 8 * ITERS / ms / 1e6,
               thr * 8 * ITERS / (ms * 1e-3) / sms / 1.965e9);
        ms = time_it([&] { k_cvt_int<<<blocks, TPB>>>(out, fin); });
        printf("ctas/sm %d  intcvt+DADD %6.3f ms  %7.2f Gcvt/s  (%.1f /clk/SM)\n", cps, ms, thr * 8 * ITERS / ms / 1e6,
               thr * 8 * ITERS / (ms * 1e-3) / sms / 1.965e9);
        ms = time_it([&] { k_rcp64h<<<blocks, TPB>>>(out, 1.5); });
        printf("ctas/sm %d  RCP64H    %8.3f ms  %7.2f Gop/s  (%.1f /clk/SM)\n", cps, ms, thr * 8 * ITERS / ms / 1e6,
               thr * 8 * ITERS / (ms * 1e-3) / sms / 1.965e9);
        cudaFree(out); cudaFree(fin);
    }

    {
        double *acc; cudaMalloc(&acc, sizeof(double) * 160 * 128 * 3);
        k_rcp_acc<<<160, 128>>>(acc);
        static double hacc[160 * 128 * 3];
        cudaMemcpy(hacc, acc, sizeof(hacc), cudaMemcpyDeviceToHost);
        double w0 = 0, w1 = 0, w2 = 0;
        for (int i = 0; i < 160 * 128; ++i) { w0 = fmax(w0, hacc[3 * i]); w1 = fmax(w1, hacc[3 * i + 1]); w2 = fmax(w2, hacc[3 * i + 2]); }
        printf("rcp.approx.ftz.f64 max rel err: seed %.3e, +1 Newton %.3e, +2 Newton %.3e\n", w0, w1, w2);
        long long *cyc, hc; cudaMalloc(&cyc, 8);
        k_dmma_lat<<<1, 32>>>(acc, 1.0000001, 1e-9, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA dependent-chain latency: %.1f cycles\n", hc / 1024.0);
        k_dfma_lat<<<1, 32>>>(acc, 1.0000001, 1e-9, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA dependent-chain latency: %.1f cycles\n", hc / 1024.0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
