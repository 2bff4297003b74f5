// fused_kernels.cu -- K1 and K2 of one ADAM step in ONE pass over the matrix ("row tiles", layout in common.cuh).
//
// The split kernels (sparse_kernels.cu) stream the matrix twice per step and round-trip w = 1/p through HBM:
// 2 x nnz x 8 B + 2 x K x m x 4 B = 4.0 GB at C3.  Here a CTA takes a tile of <= 256 consecutive rows (~1-2 k
// entries, one bulk copy), computes p and w for the rows (pAt_mul_B!, src/sparse.jl:6-21), keeps w in shared memory,
// and immediately forms the tile's contribution to g = X^T w (pAt_mulinv_B!, src/sparse.jl:25-40) by walking the same
// entries column-major through a 16-bit permutation.  What leaves the SM is one partial sum per (tile, column) --
// rows arrive sorted by genomic position (src/rnaseq_sample.jl:399-419), so a tile touches a handful of columns.
// A second, small pass adds the partials of every column in tile order.  Bytes per step: nnz x 10 B + rows x 2 B
// + partials, ~1.4 GB at C3 instead of 4.0 GB.
//
// Arithmetic = the split fast path: rows of <= 4 entries are summed in Float32 (FFMA) and inverted with
// rcp.approx, longer rows add Float32 batches of four into a Float64 sum (bit-identical p and w to k1_sell_fwd_tma);
// column sums: Float32 over <= FT_CHUNK entries, Float64 above.  No atomics, fixed orders: run-to-run identical.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "device_utils.cuh"

namespace polee {

namespace {

constexpr int FK_CONSUMER_WARPS = 8;
constexpr int FK_CONSUMERS = FK_CONSUMER_WARPS * 32;  // 256
constexpr int FK_THREADS = FK_CONSUMERS + 32;         // + producer warp
constexpr int FK_STAGES = 2;
constexpr int FK_SLOT_CAP = 512;                      // partial sums of a tile kept in shared memory

struct FkSmemHead {
    uint64_t full[FK_STAGES], empty[FK_STAGES];
    double lpsm[FK_CONSUMER_WARPS][16];
};

__host__ __device__ inline uint32_t al128(uint32_t x) { return (x + 127u) & ~127u; }

struct FkCarve {
    uint32_t stage0, blob_cap, w_tile, lrow, slots, total;
};
__host__ __device__ inline FkCarve fk_carve(uint32_t max_blob, uint32_t max_rows, uint32_t max_E, int KP) {
    FkCarve c;
    c.stage0 = al128((uint32_t)sizeof(FkSmemHead));
    c.blob_cap = al128(max_blob);
    c.w_tile = c.stage0 + FK_STAGES * c.blob_cap;
    c.lrow = c.w_tile + al128(max_rows * KP * 4u);
    c.slots = c.lrow + al128(max_E * 2u);
    c.total = c.slots + FK_SLOT_CAP * KP * 4u;
    return c;
}

template <int KP>
__device__ __forceinline__ void lds_vec(const float *p, float *v) {
    if constexpr (KP >= 4) {
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
            float4 t = reinterpret_cast<const float4 *>(p)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = p[k];
    }
}
template <int KP>
__device__ __forceinline__ void sts_vec(float *p, const float *v) {
    if constexpr (KP >= 4) {
#pragma unroll
        for (int q = 0; q < KP / 4; ++q)
            reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) p[k] = v[k];
    }
}

template <int KP, bool LP, bool WEIGHTED, bool WRITE_W>
__global__ void __launch_bounds__(FK_THREADS, 2)
    k12_fused(const FusedTileDesc *__restrict__ desc, int n_tiles, const unsigned char *__restrict__ blob,
              const float *__restrict__ xf, float *__restrict__ partial, const float *__restrict__ row_weight,
              double *__restrict__ lp_partial, float *__restrict__ w_out, float *__restrict__ gslots, uint32_t max_slots,
              uint32_t max_blob, uint32_t max_rows, uint32_t max_E) {
    extern __shared__ __align__(128) unsigned char smraw[];
    FkSmemHead &hd0 = *reinterpret_cast<FkSmemHead *>(smraw);
    const FkCarve cv = fk_carve(max_blob, max_rows, max_E, KP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FK_STAGES; ++s) {
            mbar_init(&hd0.full[s], 1);
            mbar_init(&hd0.empty[s], FK_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == FK_CONSUMER_WARPS) {
        // ------------------------------ producer: one bulk copy per tile
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const FusedTileDesc d = desc[tile];
                const int stage = it % FK_STAGES;
                mbar_wait(&hd0.empty[stage], ((it / FK_STAGES) & 1) ^ 1);
                mbar_expect_tx(&hd0.full[stage], d.bytes);
                bulk_g2s(smraw + cv.stage0 + stage * cv.blob_cap, blob + d.off, d.bytes, &hd0.full[stage]);
            }
        }
        return;
    }

    float *w_tile = reinterpret_cast<float *>(smraw + cv.w_tile);
    uint16_t *lrow_s = reinterpret_cast<uint16_t *>(smraw + cv.lrow);
    float *sm_slots = reinterpret_cast<float *>(smraw + cv.slots);
    const uint32_t tid = threadIdx.x;  // 0..255
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it % FK_STAGES;
        mbar_wait(&hd0.full[stage], (it / FK_STAGES) & 1);
        const unsigned char *b = smraw + cv.stage0 + stage * cv.blob_cap;
        const FusedHdr hd = *reinterpret_cast<const FusedHdr *>(b);
        const BlobLayout L = blob_layout(hd.rows, hd.E, hd.C, hd.chunks);
        const uint16_t *rowoff = reinterpret_cast<const uint16_t *>(b + L.rowoff);
        const float *val = reinterpret_cast<const float *>(b + L.val);
        const uint32_t *col = reinterpret_cast<const uint32_t *>(b + L.col);
        const uint16_t *perm = reinterpret_cast<const uint16_t *>(b + L.perm);
        const uint16_t *slot0 = reinterpret_cast<const uint16_t *>(b + L.slot0);
        const uint16_t *cslot = reinterpret_cast<const uint16_t *>(b + L.cslot);
        float *slots = hd.nslots <= FK_SLOT_CAP ? sm_slots : gslots + (size_t)blockIdx.x * max_slots * KP;

        // ------------------------------ pass A: p and w of the tile's rows (thread = row)
        double lpv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) lpv[k] = 0.0;
        for (uint32_t r = tid; r < hd.rows; r += FK_CONSUMERS) {
            const uint32_t e0 = rowoff[r], e1 = rowoff[r + 1], len = e1 - e0;
            double acc[KP];
            float facc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                acc[k] = 0.0;
                facc[k] = 0.0f;
            }
            for (uint32_t t0 = 0; t0 < len; t0 += 4) {
                float xv[4][KP], v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = t0 + u < len;
                    v[u] = ok ? val[e0 + t0 + u] : 0.0f;
                    if (ok) {
                        Vec<KP>::ld(xf + (size_t)col[e0 + t0 + u] * KP, xv[u]);
                        lrow_s[e0 + t0 + u] = (uint16_t)r;
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; ++k) xv[u][k] = 0.0f;
                    }
                }
#pragma unroll
                for (int k = 0; k < KP; ++k) facc[k] = v[0] * xv[0][k];
#pragma unroll
                for (int u = 1; u < 4; ++u)
#pragma unroll
                    for (int k = 0; k < KP; ++k) facc[k] = fmaf(v[u], xv[u][k], facc[k]);
                if (len > 4) {
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] += (double)facc[k];
                }
            }
            float wt = 1.0f;
            if (WEIGHTED) wt = row_weight[hd.row0 + r];
            float wv[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float rc = len > 4 ? __frcp_rn((float)acc[k]) : rcp_approx(facc[k]);
                wv[k] = WEIGHTED ? rc * wt : rc;
                if (LP) {
                    const double lg = len > 4 ? log(acc[k]) : log((double)facc[k]);
                    lpv[k] += WEIGHTED ? lg * (double)wt : lg;
                }
            }
            sts_vec<KP>(w_tile + (size_t)r * KP, wv);
            if (WRITE_W) Vec<KP>::st(w_out + (size_t)(hd.row0 + r) * KP, wv);
        }
        if (LP) {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                double v = lpv[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) hd0.lpsm[warp][k] = v;
            }
        }
        consumer_bar_sync();
        if (LP && tid < KP) {
            double s = 0.0;
            for (int wi = 0; wi < FK_CONSUMER_WARPS; ++wi) s += hd0.lpsm[wi][tid];
            lp_partial[(size_t)tile * KP + tid] = s;
        }

        // ------------------------------ pass B: column-major walk, one partial sum per (chunk, column) run
        for (uint32_t c = tid; c < hd.chunks; c += FK_CONSUMERS) {
            uint32_t slot = slot0[c];
            const uint32_t q0 = c * FT_CHUNK, q1 = min(hd.E, q0 + FT_CHUNK);
            float acc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = 0.0f;
            for (uint32_t q = q0; q < q1; ++q) {
                const uint32_t pe = perm[q], e = pe & 0x7fffu;
                const float v = val[e];
                float wv[KP];
                lds_vec<KP>(w_tile + (size_t)lrow_s[e] * KP, wv);
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] = fmaf(v, wv[k], acc[k]);
                if (pe & 0x8000u) {
                    if (hd.nslots <= FK_SLOT_CAP) {
                        sts_vec<KP>(slots + (size_t)slot * KP, acc);
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; ++k) slots[(size_t)slot * KP + k] = acc[k];
                    }
                    ++slot;
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] = 0.0f;
                }
            }
        }
        consumer_bar_sync();

        // ------------------------------ pass C: one (tile, column) partial per distinct column
        for (uint32_t j = tid / KP; j < hd.C; j += FK_CONSUMERS / KP) {
            const int k = tid % KP;
            const uint32_t s0 = cslot[j], s1 = cslot[j + 1];
            double a = 0.0;
            for (uint32_t s = s0; s < s1; ++s) a += (double)slots[(size_t)s * KP + k];
            partial[(size_t)(hd.part0 + j) * KP + k] = (float)a;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&hd0.empty[stage]);
    }
}

// g[col] (or a level-2 slot) = sum of <= FT_UNIT (tile, column) partials, in tile order.  One warp per unit.
template <int KP>
__global__ void __launch_bounds__(256)
    k_fused_combine1(const FusedUnit *__restrict__ units, int n_units, const uint32_t *__restrict__ plist,
                     const float *__restrict__ partial, double *__restrict__ g, double *__restrict__ lvl2) {
    const int unit = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (unit >= n_units) return;
    const int lane = threadIdx.x & 31;
    constexpr int G = 32 / KP;
    const int grp = lane / KP, k = lane % KP;
    const FusedUnit u = units[unit];
    double a = 0.0;
    for (uint32_t i = u.begin + grp; i < u.end; i += G) a += (double)partial[(size_t)plist[i] * KP + k];
#pragma unroll
    for (int o = 16; o >= KP; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (grp == 0) {
        if (u.out < 0)
            g[(size_t)u.col * KP + k] = a;
        else
            lvl2[(size_t)u.out * KP + k] = a;
    }
}

template <int KP>
__global__ void __launch_bounds__(256)
    k_fused_combine2(const FusedMulti *__restrict__ multi, int n_multi, const double *__restrict__ lvl2, double *__restrict__ g) {
    const int idx = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (idx >= n_multi) return;
    const int lane = threadIdx.x & 31;
    constexpr int G = 32 / KP;
    const int grp = lane / KP, k = lane % KP;
    const FusedMulti mc = multi[idx];
    double a = 0.0;
    for (uint32_t i = grp; i < mc.count; i += G) a += lvl2[(size_t)(mc.first + i) * KP + k];
#pragma unroll
    for (int o = 16; o >= KP; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (grp == 0) g[(size_t)mc.col * KP + k] = a;
}

template <int KP>
int launch_fused_t(polee_handle *h, const float *x, double *g, bool want_lp, double *lp_partial, float *w_out) {
    const FkCarve cv = fk_carve(h->ft_max_blob, h->ft_max_rows, h->ft_max_E, KP);
    const bool weighted = h->ft_row_weight != nullptr;
#define FK_LAUNCH(LPF, WF, WW)                                                                                               \
    do {                                                                                                                     \
        auto kern = k12_fused<KP, LPF, WF, WW>;                                                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv.total);              \
        if (e != cudaSuccess) return h->fail(POLEE_ECUDA, std::string("fused kernel smem: ") + cudaGetErrorString(e));       \
        kern<<<h->ft_grid, FK_THREADS, cv.total, h->stream>>>(h->ft_desc, h->ft_tiles, h->ft_blob, x, h->ft_partial,         \
                                                              h->ft_row_weight, lp_partial, w_out, h->ft_gslots,             \
                                                              h->ft_max_slots, h->ft_max_blob, h->ft_max_rows, h->ft_max_E); \
    } while (0)
    if (w_out) {
        if (weighted) FK_LAUNCH(false, true, true); else FK_LAUNCH(false, false, true);
    } else if (want_lp) {
        if (weighted) FK_LAUNCH(true, true, false); else FK_LAUNCH(true, false, false);
    } else {
        if (weighted) FK_LAUNCH(false, true, false); else FK_LAUNCH(false, false, false);
    }
#undef FK_LAUNCH
    if (h->ft_nunits > 0) {
        const int blocks = (h->ft_nunits + 7) / 8;
        k_fused_combine1<KP><<<blocks, 256, 0, h->stream>>>(h->ft_units, h->ft_nunits, h->ft_plist, h->ft_partial, g, h->ft_lvl2);
    }
    if (h->ft_nmulti > 0) {
        const int blocks = (h->ft_nmulti + 7) / 8;
        k_fused_combine2<KP><<<blocks, 256, 0, h->stream>>>(h->ft_multi, h->ft_nmulti, h->ft_lvl2, g);
    }
    return POLEE_OK;
}

}  // namespace

// CTAs for the persistent fused kernel; negative when some tile's partial sums do not fit in shared memory and the
// per-CTA global spill area is needed
int fused_grid(polee_handle *h, int KP) {
    const FkCarve cv = fk_carve(h->ft_max_blob, h->ft_max_rows, h->ft_max_E, KP);
    int per_sm = (int)std::max<uint32_t>(1, std::min<uint32_t>(2, (227u * 1024u) / (cv.total + 1024u)));
    if (const char *e = getenv("POLEE_FUSED_CTAS")) per_sm = std::max(1, atoi(e));
    const int grid = std::max(1, std::min(h->ft_tiles, h->num_sms * per_sm));
    return h->ft_max_slots > (uint32_t)FK_SLOT_CAP ? -grid : grid;
}

int launch_fused(polee_handle *h, const float *x, double *g, bool want_lp, double *lp_partial, float *w_out, int KP) {
    switch (KP) {
        case 1: return launch_fused_t<1>(h, x, g, want_lp, lp_partial, w_out);
        case 2: return launch_fused_t<2>(h, x, g, want_lp, lp_partial, w_out);
        case 4: return launch_fused_t<4>(h, x, g, want_lp, lp_partial, w_out);
        case 8: return launch_fused_t<8>(h, x, g, want_lp, lp_partial, w_out);
        case 16: return launch_fused_t<16>(h, x, g, want_lp, lp_partial, w_out);
        default: return h->fail(POLEE_EINVAL, "unsupported number of MC draws (1..16)");
    }
}

}  // namespace polee
