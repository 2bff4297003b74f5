// peer_allreduce.cu -- the per-step sum of the transcript-length gradient g[n][KP] over the ranks of a row-partitioned
// fit, as ONE kernel over NVLink peer memory instead of narrow -> ncclAllReduce -> widen.
//
// Every rank owns a block of device memory that all ranks of the box map (CUDA IPC): [send | recv | flags].  The
// kernel runs the same grid on every rank, and CTA c of a rank only ever touches "its" tiles (tile t belongs to CTA
// t mod G), so CTAs pair up across ranks and no grid-wide barrier is needed:
//   1. narrow: own tiles of the local Float64 g -> Float32 `send`;                  signal CTA c of every rank
//   2. wait for CTA c of every rank;  reduce-scatter: for the tiles of MY slice of g, add the P ranks' `send` values
//      (read over NVLink, fixed rank order, Float64 accumulator -> the same bits whoever computes them), round once to
//      Float32 and store the sum into EVERY rank's `recv` (all-gather by remote stores);  signal CTA c of every rank
//   3. wait for CTA c of every rank;  widen: own tiles of `recv` -> the local Float64 g.
// Flags carry an epoch that each CTA counts in device memory, so the captured step graph replays unchanged.  Waits time
// out (a rank that died must not hang the others' GPUs): the error code then surfaces through polee_sync.
// Nothing here has a counterpart in the reference (single process); what it must preserve is that all ranks see
// bit-identical g, because every rank runs the tree backward / ADAM step redundantly on it.
#include <cstring>
#include <mutex>
#include <shared_mutex>

#include "common.cuh"

namespace polee {

namespace {

constexpr int PEER_THREADS = 512;
constexpr int PEER_MAX_CTAS = 128;  // co-resident with room to spare (148 SMs): the CTAs spin on each other
constexpr long long PEER_TIMEOUT_CYCLES = 1ll << 32;  // ~2 s

struct PeerArgs {
    float *base[PEER_MAX_RANKS];  // every rank's block in this process's address space ([rank] = the local one)
    int P, r, G;
    size_t cap;    // floats in `send` (and in `recv`)
    size_t count;  // floats to reduce
    double *g;     // local gradient, in and out
    int *err;      // set to PEER_ERR_TIMEOUT when a wait times out
};

__device__ __forceinline__ uint32_t *flag_base(float *base, size_t cap) { return reinterpret_cast<uint32_t *>(base + 2 * cap); }

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all P ranks' CTA c have reached epoch `ep` in flag array `which` (0: data ready, 1: sums stored)
__device__ __forceinline__ void signal_and_wait(const PeerArgs &A, int which, uint32_t ep) {
    __syncthreads();  // the CTA's own loads / stores of the phase before
    const int c = blockIdx.x;
    if ((int)threadIdx.x < A.P) {
        __threadfence_system();
        uint32_t *remote = flag_base(A.base[threadIdx.x], A.cap) + ((size_t)which * PEER_MAX_RANKS + A.r) * PEER_MAX_CTAS + c;
        st_release_sys(remote, ep);
        const uint32_t *mine = flag_base(A.base[A.r], A.cap) + ((size_t)which * PEER_MAX_RANKS + threadIdx.x) * PEER_MAX_CTAS + c;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - ep) < 0) {
            if (clock64() - t0 > PEER_TIMEOUT_CYCLES) {
                atomicExch(A.err, PEER_ERR_TIMEOUT);
                break;
            }
        }
    }
    __syncthreads();
}

template <int V>  // V floats per thread per tile (4: float4 / double4-as-2x-double2, 1: scalar)
__global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce(const PeerArgs A) {
    const int c = blockIdx.x, G = A.G;
    float *send = A.base[A.r], *recv = A.base[A.r] + A.cap;
    uint32_t *epoch = flag_base(A.base[A.r], A.cap) + (size_t)2 * PEER_MAX_RANKS * PEER_MAX_CTAS + c;
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) {
        ep_s = *epoch + 1u;
        *epoch = ep_s;
    }
    __syncthreads();
    const uint32_t ep = ep_s;
    const size_t nvec = A.count / V, tile = PEER_THREADS;  // vectors; a tile = PEER_THREADS vectors
    const size_t ntiles = (nvec + tile - 1) / tile;
    // slice s of the vectors belongs to rank s
    const size_t per = ((ntiles + A.P - 1) / A.P) * tile;  // vectors per slice, whole tiles

    // Every phase works on batches of UB tiles: all loads of a batch are issued before its first store (the loops carry no
    // dependence, but the compiler cannot know that `g`, `send` and `recv` never overlap)
    constexpr int UB = 4;

    // ---- 1. narrow my tiles (of every slice)
    for (size_t t = c; t < ntiles; t += (size_t)G * UB) {
        if constexpr (V == 4) {
            double2 a[UB], b[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) {
                    a[u] = reinterpret_cast<const double2 *>(A.g)[2 * v];
                    b[u] = reinterpret_cast<const double2 *>(A.g)[2 * v + 1];
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) reinterpret_cast<float4 *>(send)[v] = make_float4((float)a[u].x, (float)a[u].y, (float)b[u].x, (float)b[u].y);
            }
        } else {
            double a[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) a[u] = A.g[v];
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) send[v] = (float)a[u];
            }
        }
    }
    signal_and_wait(A, 0, ep);

    // ---- 2. my slice: sum over the ranks, store everywhere
    {
        const size_t t0 = (size_t)A.r * (per / tile), v1 = (t0 + per / tile) * tile < nvec ? (t0 + per / tile) * tile : nvec;
        // the tiles of my slice that are CTA c's everywhere (t mod G == c): the flags pair CTA c with CTA c only
        for (size_t t = t0 + ((size_t)c + G - t0 % G) % G; t * tile < v1; t += G) {
            const size_t v = t * tile + threadIdx.x;
            if (v >= v1) continue;
            if constexpr (V == 4) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                float4 in[PEER_MAX_RANKS];
#pragma unroll
                for (int p = 0; p < PEER_MAX_RANKS; ++p)
                    if (p < A.P) in[p] = __ldcg(reinterpret_cast<const float4 *>(A.base[p]) + v);
#pragma unroll
                for (int p = 0; p < PEER_MAX_RANKS; ++p)
                    if (p < A.P) {
                        s0 += (double)in[p].x; s1 += (double)in[p].y; s2 += (double)in[p].z; s3 += (double)in[p].w;
                    }
                const float4 out = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
#pragma unroll
                for (int p = 0; p < PEER_MAX_RANKS; ++p)
                    if (p < A.P) reinterpret_cast<float4 *>(A.base[p] + A.cap)[v] = out;
            } else {
                double s = 0.0;
                for (int p = 0; p < A.P; ++p) s += (double)__ldcg(A.base[p] + v);
                for (int p = 0; p < A.P; ++p) (A.base[p] + A.cap)[v] = (float)s;
            }
        }
    }
    signal_and_wait(A, 1, ep);

    // ---- 3. widen my tiles (of every slice)
    for (size_t t = c; t < ntiles; t += (size_t)G * UB) {
        if constexpr (V == 4) {
            float4 a[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) a[u] = __ldcg(reinterpret_cast<const float4 *>(recv) + v);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) {
                    reinterpret_cast<double2 *>(A.g)[2 * v] = make_double2((double)a[u].x, (double)a[u].y);
                    reinterpret_cast<double2 *>(A.g)[2 * v + 1] = make_double2((double)a[u].z, (double)a[u].w);
                }
            }
        } else {
            float a[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) a[u] = __ldcg(recv + v);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const size_t v = (t + (size_t)u * G) * tile + threadIdx.x;
                if (v < nvec) A.g[v] = (double)a[u];
            }
        }
    }
}

size_t peer_block_bytes(size_t cap) { return 2 * cap * sizeof(float) + (size_t)(2 * PEER_MAX_RANKS + 1) * PEER_MAX_CTAS * sizeof(uint32_t); }

}  // namespace

void peer_release(polee_handle *h) {
    for (int p = 0; p < PEER_MAX_RANKS; ++p) {
        if (h->peer_base[p] && p != h->rank) cudaIpcCloseMemHandle(h->peer_base[p]);
        h->peer_base[p] = nullptr;
    }
    if (h->peer_local) cudaFree(h->peer_local);
    h->peer_local = nullptr;
    h->peer_ready = false;
    h->peer_cap = 0;
}

int launch_peer_allreduce(polee_handle *h, double *g, size_t count) {
    PeerArgs A;
    for (int p = 0; p < PEER_MAX_RANKS; ++p) A.base[p] = reinterpret_cast<float *>(h->peer_base[p]);
    A.P = h->nranks; A.r = h->rank; A.cap = h->peer_cap; A.count = count; A.g = g; A.err = h->d_bad_step;
    if (count > h->peer_cap) return h->fail(POLEE_EINVAL, "peer all-reduce: buffer smaller than the gradient");
    const bool vec = count % 4 == 0;
    const size_t nvec = vec ? count / 4 : count, ntiles = (nvec + PEER_THREADS - 1) / PEER_THREADS;
    A.G = (int)std::max<size_t>(1, std::min<size_t>(PEER_MAX_CTAS, ntiles));
    if (vec) k_peer_allreduce<4><<<A.G, PEER_THREADS, 0, h->stream>>>(A);
    else k_peer_allreduce<1><<<A.G, PEER_THREADS, 0, h->stream>>>(A);
    return POLEE_OK;
}

}  // namespace polee

#define CK(expr) POLEE_CUDA_CHECK(h, expr)

// Allocate this rank's block and hand out its CUDA IPC handle (64 bytes); the host exchanges the handles of all ranks
// (any transport: torch.distributed in bench.py, MPI, a file ...) and passes them to polee_comm_peer_import.
extern "C" int polee_comm_peer_export(polee_handle *h, char handle[64]) {
    if (!h) return POLEE_EINVAL;
    h->err.clear();
    if (cudaSetDevice(h->device) != cudaSuccess) return h->fail(POLEE_ECUDA, "cudaSetDevice failed");
    if (!handle) return h->fail(POLEE_EINVAL, "comm_peer_export: null pointer");
    if (h->nranks < 2 || h->nranks > polee::PEER_MAX_RANKS) return h->fail(POLEE_EINVAL, "comm_peer_export: call polee_comm_init with 2..16 ranks first");
    const int64_t n = h->have_matrix ? h->n : (h->have_tree ? h->td.n : 0);
    if (n < 1) return h->fail(POLEE_EINVAL, "comm_peer_export: set the matrix or the tree first (n unknown)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    polee::drop_step_graph(h);
    polee::peer_release(h);
    const size_t cap = (((size_t)(n + 1) * 16) + 3) & ~(size_t)3;  // room for KP up to 16
    {
        std::unique_lock<std::shared_mutex> cap_lock(polee::capture_mutex());
        CK(cudaMalloc(&h->peer_local, polee::peer_block_bytes(cap)));  // IPC needs a whole cudaMalloc allocation
        CK(cudaMemset(h->peer_local, 0, polee::peer_block_bytes(cap)));
        CK(cudaDeviceSynchronize());
    }
    h->peer_cap = cap;
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, h->peer_local));
    std::memcpy(handle, &ih, 64);
    return POLEE_OK;
}

extern "C" int polee_comm_peer_import(polee_handle *h, const char *handles) {
    if (!h) return POLEE_EINVAL;
    h->err.clear();
    if (cudaSetDevice(h->device) != cudaSuccess) return h->fail(POLEE_ECUDA, "cudaSetDevice failed");
    polee::drop_step_graph(h);
    if (!handles) {  // back to the NCCL all-reduce
        polee::peer_release(h);
        return POLEE_OK;
    }
    if (!h->peer_local) return h->fail(POLEE_EINVAL, "comm_peer_import: call polee_comm_peer_export first");
    for (int p = 0; p < h->nranks; ++p) {
        if (p == h->rank) {
            h->peer_base[p] = h->peer_local;
            continue;
        }
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, handles + (size_t)p * 64, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            polee::peer_release(h);
            return h->fail(POLEE_ECUDA, std::string("comm_peer_import: cudaIpcOpenMemHandle (rank ") + std::to_string(p) +
                                            "): " + cudaGetErrorString(e) + " -- the NCCL all-reduce stays in use");
        }
        h->peer_base[p] = ptr;
    }
    h->peer_ready = true;
    return POLEE_OK;
}
