// sparse_kernels.cu -- the two HBM-bound sparse passes of one ADAM step, for KP (= K padded to a
// power of two) Monte-Carlo draws at once, so the matrix is streamed once per step instead of once
// per draw as in the reference.
//
//  K1  fragment likelihood   p[i][k] = sum_j X[i][j] * x[j][k]       pAt_mul_B!(frag_probs, Xt, xs)
//      fused with w = 1/p (or ks/p) and, optionally, sum_i log p      src/sparse.jl:6-21, likelihood.jl:21-25,46-51,78-80
//  K2  transposed gradient   g[j][k] = sum_i X[i][j] * w[i][k]       pAt_mulinv_B!(x_grad, X, frag_probs)
//      deterministic: no atomics, fixed reduction trees              src/sparse.jl:25-40
//
// Arithmetic contract: each product x*v is rounded to Float32 and summed in
// Float64 in ascending-transcript order, exactly as the reference does -> p is bit-identical to the
// reference's frag_probs.  w is stored as Float32 (one rounding, <= 2^-24 relative); g accumulates
// Float64(v) * Float64(w) in Float64.  Tensor cores are not used: nothing here is a dense contraction.
#include "common.cuh"

namespace polee {

namespace {

// ---------------------------------------------------------------- load / store helpers
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int KP>
struct Vec;
template <>
struct Vec<1> {
    static __device__ __forceinline__ void ld(const float *p, float *v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float *p, const float *v) { p[0] = v[0]; }
};
template <>
struct Vec<2> {
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float *p, const float *v) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};
template <>
struct Vec<4> {
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float *p, const float *v) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Vec<8> {
    // one 256-bit request per lane (sm_100: LDG.E.ENL2.256): a whole 32-byte sector per entry
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    }
    static __device__ __forceinline__ void st(float *p, const float *v) {
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                     "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                     : "memory");
    }
};
template <>
struct Vec<16> {
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        Vec<8>::ld(p, v);
        Vec<8>::ld(p + 8, v + 8);
    }
    static __device__ __forceinline__ void st(float *p, const float *v) {
        Vec<8>::st(p, v);
        Vec<8>::st(p + 8, v + 8);
    }
};

// ---------------------------------------------------------------- K1
// One thread = one row (fragment) of one row-length class; KP Float64 accumulators in registers.
template <int KP, bool LP, bool WEIGHTED>
__global__ void __launch_bounds__(ROW_TILE)
    k1_sell_fwd(const RowTile *__restrict__ tiles, const uint32_t *__restrict__ idx, const float *__restrict__ val,
                const float *__restrict__ x, float *__restrict__ w, const float *__restrict__ row_weight,
                double *__restrict__ lp_partial) {
    const RowTile t = tiles[blockIdx.x];
    const uint32_t r = threadIdx.x;
    double lpv[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) lpv[k] = 0.0;

    if (r < t.nrows) {
        const uint32_t *ip = idx + t.slab_off + r;
        const float *vp = val + t.slab_off + r;
        const size_t stride = t.stride;
        const uint32_t L = t.len;
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;
        uint32_t tt = 0;
        constexpr int U = (KP >= 16) ? 2 : 4;
        for (; tt + U <= L; tt += U) {
            uint32_t c[U];
            float v[U];
            float xv[U][KP];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                c[u] = ld_stream_u32(ip + (size_t)(tt + u) * stride);
                v[u] = ld_stream_f32(vp + (size_t)(tt + u) * stride);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) Vec<KP>::ld(x + (size_t)c[u] * KP, xv[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] = __dadd_rn(acc[k], (double)__fmul_rn(xv[u][k], v[u]));
        }
        for (; tt < L; ++tt) {
            uint32_t c = ld_stream_u32(ip + (size_t)tt * stride);
            float v = ld_stream_f32(vp + (size_t)tt * stride);
            float xv[KP];
            Vec<KP>::ld(x + (size_t)c * KP, xv);
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = __dadd_rn(acc[k], (double)__fmul_rn(xv[k], v));
        }
        float wt = 1.0f;
        if (WEIGHTED) wt = row_weight[t.row0 + r];
        float wv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            float rc = __frcp_rn((float)acc[k]);
            wv[k] = WEIGHTED ? rc * wt : rc;
        }
        Vec<KP>::st(w + (size_t)(t.row0 + r) * KP, wv);
        if (LP) {
#pragma unroll
            for (int k = 0; k < KP; ++k) lpv[k] = WEIGHTED ? log(acc[k]) * (double)wt : log(acc[k]);
        }
    }
    if (LP) {
        // fixed-order block reduction: shuffle tree inside each warp, then warp 0 adds the 8 warp sums
        __shared__ double sm[ROW_TILE / 32][KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            double v = lpv[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < KP) {
            double s = 0.0;
            for (int wi = 0; wi < ROW_TILE / 32; ++wi) s += sm[wi][threadIdx.x];
            lp_partial[(size_t)blockIdx.x * KP + threadIdx.x] = s;
        }
    }
}

// ---------------------------------------------------------------- K2
// Fixed butterfly: after it, the lane holds in acc[0] the warp total of draw k = draw_of_lane<KP>(lane).
template <int KP>
__device__ __forceinline__ void warp_reduce_scatter(double (&acc)[KP], int lane) {
    int mask = 16;
#pragma unroll
    for (int half = KP / 2; half >= 1; half >>= 1, mask >>= 1) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            double keep = up ? acc[i + half] : acc[i];
            double send = up ? acc[i] : acc[i + half];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    for (; mask >= 1; mask >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], mask);
}
template <int KP>
__device__ __forceinline__ int draw_of_lane(int lane) {
    // bits consumed from the top: lane bit 4 is the most significant bit of k
    int k = 0, mask = 16;
    for (int half = KP / 2; half >= 1; half >>= 1, mask >>= 1) k = (k << 1) | ((lane & mask) ? 1 : 0);
    return k;
}
template <int KP>
__device__ __forceinline__ bool lane_writes(int lane) {
    // one writer per draw: the lanes whose unconsumed low bits are zero
    int consumed = 0;
    for (int half = KP / 2; half >= 1; half >>= 1) ++consumed;
    int low_mask = (1 << (5 - consumed)) - 1;
    return (lane & low_mask) == 0;
}

constexpr int K2_WARPS = 8;

// One warp = one column segment (<= COL_SEG consecutive entries of one column, rows ascending):
// lane l takes entries l, l+32, ... so a warp-wide gather touches neighbouring rows of w.
template <int KP>
__global__ void __launch_bounds__(K2_WARPS * 32)
    k2_csc_grad(const ColSeg *__restrict__ segs, int n_segs, const uint32_t *__restrict__ csc_row,
                const float *__restrict__ csc_val, const float *__restrict__ w, double *__restrict__ g,
                double *__restrict__ seg_partial) {
    const int lane = threadIdx.x & 31;
    const int sidx = blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
    if (sidx >= n_segs) return;
    const ColSeg sg = segs[sidx];
    double acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.0;
    const uint32_t *rp = csc_row + sg.start;
    const float *vp = csc_val + sg.start;
    constexpr int U = (KP >= 16) ? 2 : 4;
    uint32_t off = lane;
    for (; off + 32 * (U - 1) < sg.len; off += 32 * U) {
        uint32_t r[U];
        float v[U];
        float wv[U][KP];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            r[u] = ld_stream_u32(rp + off + 32 * u);
            v[u] = ld_stream_f32(vp + off + 32 * u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) Vec<KP>::ld(w + (size_t)r[u] * KP, wv[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double dv = (double)v[u];
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = fma(dv, (double)wv[u][k], acc[k]);
        }
    }
    for (; off < sg.len; off += 32) {
        uint32_t r = ld_stream_u32(rp + off);
        const double dv = (double)ld_stream_f32(vp + off);
        float wv[KP];
        Vec<KP>::ld(w + (size_t)r * KP, wv);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = fma(dv, (double)wv[k], acc[k]);
    }
    warp_reduce_scatter<KP>(acc, lane);
    if (lane_writes<KP>(lane)) {
        const int k = draw_of_lane<KP>(lane);
        if (sg.slot < 0)
            g[(size_t)sg.col * KP + k] = acc[0];
        else
            seg_partial[(size_t)sg.slot * KP + k] = acc[0];
    }
}

// columns that span several segments: add their partials in segment order
__global__ void k2_combine(const MultiCol *__restrict__ multi, int n_multi, int KP,
                           const double *__restrict__ seg_partial, double *__restrict__ g) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_multi * KP) return;
    const MultiCol mc = multi[t / KP];
    const int k = t % KP;
    double s = 0.0;
    for (uint32_t q = 0; q < mc.nslots; ++q) s += seg_partial[(size_t)(mc.first_slot + q) * KP + k];
    g[(size_t)mc.col * KP + k] = s;
}

// single CTA, fixed order: out[k] = sum_t partial[t][k]
__global__ void __launch_bounds__(1024) k_reduce_partials(const double *__restrict__ partial, int count, int KP,
                                                          double *__restrict__ out) {
    __shared__ double sm[1024];
    const int k = threadIdx.x % KP, lane_t = threadIdx.x / KP, per = 1024 / KP;
    double s = 0.0;
    for (int t = lane_t; t < count; t += per) s += partial[(size_t)t * KP + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int span = per / 2; span >= 1; span >>= 1) {
        if (lane_t < span) sm[threadIdx.x] += sm[threadIdx.x + span * KP];
        __syncthreads();
    }
    if (threadIdx.x < KP) out[threadIdx.x] = sm[threadIdx.x];
}

template <int KP>
int launch_k1_t(polee_handle *h, const float *x, float *w, bool want_lp, double *lp_partial) {
    if (h->n_row_tiles == 0) return POLEE_OK;
    dim3 grid(h->n_row_tiles), block(ROW_TILE);
    const bool wt = h->row_weight != nullptr;
#define K1_LAUNCH(LPF, WF)                                                                                      \
    k1_sell_fwd<KP, LPF, WF><<<grid, block, 0, h->stream>>>(h->row_tiles, h->sell_idx, h->sell_val, x, w,       \
                                                            h->row_weight, lp_partial)
    if (want_lp) {
        if (wt) K1_LAUNCH(true, true); else K1_LAUNCH(true, false);
    } else {
        if (wt) K1_LAUNCH(false, true); else K1_LAUNCH(false, false);
    }
#undef K1_LAUNCH
    return POLEE_OK;
}

template <int KP>
int launch_k2_t(polee_handle *h, const float *w, double *g) {
    if (h->n_segs > 0) {
        dim3 grid((h->n_segs + K2_WARPS - 1) / K2_WARPS), block(K2_WARPS * 32);
        k2_csc_grad<KP><<<grid, block, 0, h->stream>>>(h->segs, h->n_segs, h->csc_row, h->csc_val, w, g,
                                                       h->seg_partial);
    }
    if (h->n_multi > 0) {
        int work = h->n_multi * KP;
        k2_combine<<<(work + 255) / 256, 256, 0, h->stream>>>(h->multi, h->n_multi, KP, h->seg_partial, g);
    }
    return POLEE_OK;
}

}  // namespace

#define DISPATCH_KP(KP, CALL)                                                   \
    switch (KP) {                                                               \
        case 1: { constexpr int KPC = 1; CALL; } break;                         \
        case 2: { constexpr int KPC = 2; CALL; } break;                         \
        case 4: { constexpr int KPC = 4; CALL; } break;                         \
        case 8: { constexpr int KPC = 8; CALL; } break;                         \
        case 16: { constexpr int KPC = 16; CALL; } break;                       \
        default: return h->fail(POLEE_EINVAL, "unsupported number of MC draws (1..16)"); \
    }

int launch_k1(polee_handle *h, const float *x, float *w, bool want_lp, double *lp_partial, int KP) {
    int rc = POLEE_OK;
    DISPATCH_KP(KP, rc = launch_k1_t<KPC>(h, x, w, want_lp, lp_partial));
    return rc;
}

int launch_k2(polee_handle *h, const float *w, double *g, int KP) {
    int rc = POLEE_OK;
    DISPATCH_KP(KP, rc = launch_k2_t<KPC>(h, w, g));
    return rc;
}

int launch_reduce_lp(polee_handle *h, const double *lp_partial, double *lp, int KP) {
    k_reduce_partials<<<1, 1024, 0, h->stream>>>(lp_partial, h->n_row_tiles, KP, lp);
    return POLEE_OK;
}

}  // namespace polee
