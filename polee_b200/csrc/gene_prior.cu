// gene_prior.cu -- gene_noninformative_prior! (src/likelihood.jl:114-159) on the device.
//
// The optional `--gene-noninformative` prior of the reference (off by default, src/main.jl:712) adds, per MC draw,
//   xl_grad_i = -(k_g - 1) / c_g        for every transcript i of a gene g with k_g > 1 transcripts,
//                                       c_g = sum_{i in g} xls_i,  xls_i = Float32(Float32(x_i / l_i) / S)   (:118-133, :98-101)
//   offdiag   = (sum_i -xl_grad_i xls_i) / S^2                                                               (:143-147)
//   x_grad_i += xl_grad_i ((1/l_i) / S) + (1/l_i) offdiag                                                    (:149-154)
// with S = sum_i x_i / l_i (the same Float64 sum the effective-length adjustment uses, likelihood.jl:96-100).
// Here the K draws of a step are processed together ([item][KP] layouts) and the term is added to the all-reduced
// g right before the tree backward pass; sums run in a fixed order (no atomics), so results are run-to-run identical.
#include "common.cuh"

namespace polee {

namespace {

constexpr int GP_THREADS = 256;

// one thread per (gene, draw); genes here are only those with >= 2 transcripts
template <int KP>
__global__ void __launch_bounds__(GP_THREADS)
    k_gene_sums(int64_t n_genes, const int64_t *__restrict__ gene_ptr, const int32_t *__restrict__ gene_tx,
                const float *__restrict__ x, const float *__restrict__ efflen, const double *__restrict__ S,
                double *__restrict__ xl_grad, double *__restrict__ off_partial) {
    __shared__ double sm[GP_THREADS];
    const int64_t tid = (int64_t)blockIdx.x * GP_THREADS + threadIdx.x;
    const int64_t gene = tid / KP;
    const int k = (int)(tid % KP);
    double off = 0.0;
    if (gene < n_genes) {
        const int64_t b = gene_ptr[gene], e = gene_ptr[gene + 1];
        const double Sk = S[k];
        double c = 0.0;
        for (int64_t q = b; q < e; ++q) {
            const int i = gene_tx[q];
            const float xl = (float)((double)__fdiv_rn(x[(size_t)i * KP + k], efflen[i]) / Sk);  // :98, :101
            c += (double)xl;                                                                      // :124-127
        }
        const double xg = -(double)(e - b - 1) / c;  // :130
        for (int64_t q = b; q < e; ++q) {
            const int i = gene_tx[q];
            const float xl = (float)((double)__fdiv_rn(x[(size_t)i * KP + k], efflen[i]) / Sk);
            xl_grad[(size_t)i * KP + k] = xg;
            off += -xg * (double)xl;  // :145
        }
    }
    sm[threadIdx.x] = off;
    __syncthreads();
    for (int span = GP_THREADS / 2; span >= KP; span >>= 1) {  // lanes with equal k are KP apart
        if ((int)threadIdx.x < span) sm[threadIdx.x] += sm[threadIdx.x + span];
        __syncthreads();
    }
    if ((int)threadIdx.x < KP) off_partial[(size_t)blockIdx.x * KP + threadIdx.x] = sm[threadIdx.x];
}

template <int KP>
__global__ void __launch_bounds__(GP_THREADS)
    k_gene_off_reduce(int n_blocks, const double *__restrict__ off_partial, const double *__restrict__ S,
                      double *__restrict__ off) {
    __shared__ double sm[GP_THREADS];
    const int k = threadIdx.x % KP;
    double a = 0.0;
    for (int b = threadIdx.x / KP; b < n_blocks; b += GP_THREADS / KP) a += off_partial[(size_t)b * KP + k];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int span = GP_THREADS / 2; span >= KP; span >>= 1) {
        if ((int)threadIdx.x < span) sm[threadIdx.x] += sm[threadIdx.x + span];
        __syncthreads();
    }
    if ((int)threadIdx.x < KP) off[threadIdx.x] = sm[threadIdx.x] / (S[threadIdx.x] * S[threadIdx.x]);  // :141, :147
}

template <int KP>
__global__ void __launch_bounds__(GP_THREADS)
    k_gene_apply(int64_t n, const float *__restrict__ efflen, const double *__restrict__ S,
                 const double *__restrict__ xl_grad, const double *__restrict__ off, double *__restrict__ g) {
    const int64_t tid = (int64_t)blockIdx.x * GP_THREADS + threadIdx.x;
    if (tid >= n * KP) return;
    const int64_t i = tid / KP;
    const int k = (int)(tid % KP);
    const float inv_l = __fdiv_rn(1.0f, efflen[i]);                 // 1/efflens[i] is Float32
    const double grad_a = xl_grad[tid] * ((double)inv_l / S[k]);    // :150
    const double grad_b = (double)inv_l * off[k];                   // :151
    g[tid] += grad_a + grad_b;                                      // :152-153
}

}  // namespace

#define CK(expr) POLEE_CUDA_CHECK(h, expr)

void release_gene_buffers(polee_handle *h) {
    polee::dfree(h->gene_xl_grad);
    polee::dfree(h->gene_off_partial);
    polee::dfree(h->gene_off);
    h->gene_xl_grad = h->gene_off_partial = h->gene_off = nullptr;
    h->gene_KP = 0;
}

int ensure_gene_buffers(polee_handle *h, int KP) {
    if (h->n_genes == 0 || h->gene_KP == KP) return POLEE_OK;
    release_gene_buffers(h);
    const int blocks = (int)((h->n_genes * KP + GP_THREADS - 1) / GP_THREADS);
    CK(polee::dmalloc((void **)&h->gene_xl_grad, sizeof(double) * (size_t)h->n * KP));
    CK(cudaMemsetAsync(h->gene_xl_grad, 0, sizeof(double) * (size_t)h->n * KP, h->stream));  // transcripts outside multi-transcript genes: 0 (:118)
    CK(polee::dmalloc((void **)&h->gene_off_partial, sizeof(double) * (size_t)blocks * KP));
    CK(polee::dmalloc((void **)&h->gene_off, sizeof(double) * KP));
    h->gene_KP = KP;
    return POLEE_OK;
}

int launch_gene_prior(polee_handle *h, int KP) {
    if (h->n_genes == 0) return POLEE_OK;
    if (h->gene_KP != KP) return h->fail(POLEE_EINVAL, "gene prior buffers were not prepared");
    const int blocks = (int)((h->n_genes * KP + GP_THREADS - 1) / GP_THREADS);
    const int ablocks = (int)((h->n * KP + GP_THREADS - 1) / GP_THREADS);
    switch (KP) {
#define GP_CASE(KPC)                                                                                                         \
    case KPC:                                                                                                                \
        k_gene_sums<KPC><<<blocks, GP_THREADS, 0, h->stream>>>(h->n_genes, h->gene_ptr, h->gene_tx, h->x, h->efflen, h->S,   \
                                                               h->gene_xl_grad, h->gene_off_partial);                        \
        k_gene_off_reduce<KPC><<<1, GP_THREADS, 0, h->stream>>>(blocks, h->gene_off_partial, h->S, h->gene_off);             \
        k_gene_apply<KPC><<<ablocks, GP_THREADS, 0, h->stream>>>(h->n, h->efflen, h->S, h->gene_xl_grad, h->gene_off, h->g); \
        break;
        GP_CASE(1) GP_CASE(2) GP_CASE(4) GP_CASE(8) GP_CASE(16)
#undef GP_CASE
        default: return h->fail(POLEE_EINVAL, "unsupported number of MC draws (1..16)");
    }
    return POLEE_OK;
}

}  // namespace polee
