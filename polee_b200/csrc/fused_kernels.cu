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

constexpr int FW_WARPS = 4;  // warps per CTA; every warp streams its own tiles (no CTA-wide barrier anywhere)

__host__ __device__ inline uint32_t al128(uint32_t x) { return (x + 127u) & ~127u; }

struct FwCarve {
    uint32_t head, w_tile, xs, lrow, per_warp, total;  // per-warp offsets (the ring sits at 0)
};
// xs = x of the tile's columns during pass A, then the partial-sum slots of pass B (never live together)
__host__ __device__ inline FwCarve fw_carve(uint32_t ring_bytes, uint32_t max_rows, uint32_t max_E, uint32_t max_C,
                                            uint32_t max_slots, int KP) {
    FwCarve c;
    c.head = 128;  // FW_WARPS x 2 mbarriers
    c.w_tile = al128(ring_bytes);
    c.xs = c.w_tile + al128(max_rows * KP * 4u);
    c.lrow = c.xs + al128((max_C > max_slots ? max_C : max_slots) * KP * 4u);
    c.per_warp = c.lrow + al128(max_E);
    c.total = c.head + FW_WARPS * c.per_warp;
    return c;
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

// KP floats as packed pairs: Blackwell issues two Float32 FMAs per FFMA2 (fma.rn.f32x2), each exactly fmaf
template <int KP>
struct Acc {
    static constexpr int NP = (KP + 1) / 2;
    float2 v[NP];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] = make_float2(0.0f, 0.0f);
    }
    __device__ __forceinline__ void lds(const float *p) {
        if constexpr (KP >= 4) {
#pragma unroll
            for (int q = 0; q < KP / 4; ++q) {
                const float4 t = reinterpret_cast<const float4 *>(p)[q];
                v[2 * q] = make_float2(t.x, t.y);
                v[2 * q + 1] = make_float2(t.z, t.w);
            }
        } else if constexpr (KP == 2) {
            v[0] = *reinterpret_cast<const float2 *>(p);
        } else {
            v[0] = make_float2(p[0], 0.0f);
        }
    }
    __device__ __forceinline__ void sts(float *p) const {
        if constexpr (KP >= 4) {
#pragma unroll
            for (int q = 0; q < KP / 4; ++q)
                reinterpret_cast<float4 *>(p)[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        } else if constexpr (KP == 2) {
            *reinterpret_cast<float2 *>(p) = v[0];
        } else {
            p[0] = v[0].x;
        }
    }
    __device__ __forceinline__ float get(int k) const { return (k & 1) ? v[k >> 1].y : v[k >> 1].x; }
    // this += s * o   (per element fmaf(s, o, this))
    __device__ __forceinline__ void fma(float s, const Acc &o) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
                "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
                : "+f"(v[i].x), "+f"(v[i].y)
                : "f"(s), "f"(o.v[i].x), "f"(o.v[i].y));
        }
    }
};

template <int KP, bool LP, bool WEIGHTED, bool WRITE_W>
__global__ void __launch_bounds__(FW_WARPS * 32, 4)
    k12_fused(const FusedTileDesc *__restrict__ desc, int n_tiles, const unsigned char *__restrict__ blob,
              const float *__restrict__ xf, float *__restrict__ partial, const float *__restrict__ row_weight,
              const uint32_t *__restrict__ row_of_pos, double *__restrict__ lp_partial, float *__restrict__ w_out,
              uint32_t ring_bytes, uint32_t max_rows, uint32_t max_E, uint32_t max_C, uint32_t max_slots) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const FwCarve cv = fw_carve(ring_bytes, max_rows, max_E, max_C, max_slots, KP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw) + 2 * warp;
    unsigned char *base = smraw + cv.head + (size_t)warp * cv.per_warp;
    unsigned char *ring = base;
    float *w_tile = reinterpret_cast<float *>(base + cv.w_tile);
    float *xs = reinterpret_cast<float *>(base + cv.xs);
    uint8_t *lrow = base + cv.lrow;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nw = gridDim.x * FW_WARPS;
    int tile = blockIdx.x * FW_WARPS + warp;
    if (tile >= n_tiles) return;

    // ring state (warp-uniform): the current tile lives at [cur_off, cur_off + cur_bytes)
    uint32_t cur_off = 0, cur_bytes = 0, phase = 0;  // phase bit b = parity to wait for on bar[b]
    int cur_bar = 0;
    FusedTileDesc dn{0, 0, 0};  // descriptor of the next tile, fetched one tile ahead (lane 0)
    if (lane == 0) {
        const FusedTileDesc d = desc[tile];
        mbar_expect_tx(&bar[0], d.bytes);
        bulk_g2s(ring, blob + d.off, d.bytes, &bar[0]);
        cur_bytes = d.bytes;
        if (tile + nw < n_tiles) dn = desc[tile + nw];
    }
    cur_bytes = __shfl_sync(0xffffffffu, cur_bytes, 0);

    for (;;) {
        const int next = tile + nw;
        uint32_t pre = 0, noff = 0, nbytes = 0;
        if (next < n_tiles) {
            if (lane == 0) {
                nbytes = dn.bytes;
                const uint32_t cand = cur_off + cur_bytes;
                if (cand + nbytes <= ring_bytes) {
                    noff = cand;
                    pre = 1;
                } else if (nbytes <= cur_off) {
                    noff = 0;
                    pre = 1;
                }
                if (pre) {
                    mbar_expect_tx(&bar[cur_bar ^ 1], nbytes);
                    bulk_g2s(ring + noff, blob + dn.off, nbytes, &bar[cur_bar ^ 1]);
                }
            }
            pre = __shfl_sync(0xffffffffu, pre, 0);
            noff = __shfl_sync(0xffffffffu, noff, 0);
            nbytes = __shfl_sync(0xffffffffu, nbytes, 0);
        }
        const uint64_t dn_off = dn.off;  // lane 0 only
        if (lane == 0 && next + nw < n_tiles) dn = desc[next + nw];

        mbar_wait(&bar[cur_bar], (phase >> cur_bar) & 1u);
        phase ^= 1u << cur_bar;
        const unsigned char *b = ring + cur_off;
        const FusedHdr hd = *reinterpret_cast<const FusedHdr *>(b);
        const BlobLayout L = blob_layout(hd.rows, hd.E, hd.C);
        const uint32_t *cols = reinterpret_cast<const uint32_t *>(b + L.cols);
        const uint16_t *rowoff = reinterpret_cast<const uint16_t *>(b + L.rowoff);
        const float *val = reinterpret_cast<const float *>(b + L.val);
        const uint16_t *perm = reinterpret_cast<const uint16_t *>(b + L.perm);
        const uint16_t *slot0 = reinterpret_cast<const uint16_t *>(b + L.slot0);
        const uint16_t *cslot = reinterpret_cast<const uint16_t *>(b + L.cslot);
        const uint8_t *lcol = b + L.lcol;

        // ------------------------------ x of the tile's columns -> shared memory (the only gather of the tile)
        for (uint32_t j = lane; j < hd.C; j += 32) {
            float xv[KP];
            Vec<KP>::ld(xf + (size_t)cols[j] * KP, xv);
            sts_vec<KP>(xs + (size_t)j * KP, xv);
        }
        __syncwarp();

        // ------------------------------ pass A: p and w of the tile's rows (lane = row; rows sorted longest first)
        double lpv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) lpv[k] = 0.0;
        for (uint32_t r = lane; r < hd.rows; r += 32) {
            const uint32_t e0 = rowoff[r], len = rowoff[r + 1] - e0;
            Acc<KP> facc;
            facc.zero();
            double acc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = 0.0;
            for (uint32_t t = 0; t < len; ++t) {
                if ((t & 3u) == 0u && t != 0u) {  // every four products the Float32 batch is added to the Float64 row sum
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] += (double)facc.get(k);
                    facc.zero();
                }
                const float v = val[e0 + t];
                Acc<KP> xv;
                xv.lds(xs + (size_t)lcol[e0 + t] * KP);
                lrow[e0 + t] = (uint8_t)r;
                facc.fma(v, xv);
            }
            float wt = 1.0f;
            if (WEIGHTED) wt = row_weight[hd.row0 + r];
            float wv[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float rc;
                if (len > 4) {
                    acc[k] += (double)facc.get(k);
                    rc = __frcp_rn((float)acc[k]);
                    if (LP) lpv[k] += WEIGHTED ? log(acc[k]) * (double)wt : log(acc[k]);
                } else {
                    rc = rcp_approx(facc.get(k));
                    if (LP) lpv[k] += WEIGHTED ? log((double)facc.get(k)) * (double)wt : log((double)facc.get(k));
                }
                wv[k] = WEIGHTED ? rc * wt : rc;
            }
            sts_vec<KP>(w_tile + (size_t)r * KP, wv);
            if (WRITE_W) Vec<KP>::st(w_out + (size_t)row_of_pos[hd.row0 + r] * KP, wv);
        }
        if (LP) {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                double v = lpv[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == k) lp_partial[(size_t)tile * KP + k] = v;
            }
        }
        __syncwarp();

        // ------------------------------ pass B: lane walks its chunk of the column-major order, one slot per run
        {
            const uint32_t q0 = lane * hd.chunk, q1 = min(hd.E, q0 + hd.chunk);
            uint32_t slot = q0 < hd.E ? slot0[lane] : 0u;
            Acc<KP> acc;
            acc.zero();
            for (uint32_t q = q0; q < q1; ++q) {
                const uint32_t pe = perm[q], e = pe & 0x7fffu;
                const float v = val[e];
                Acc<KP> wv;
                wv.lds(w_tile + (size_t)lrow[e] * KP);
                acc.fma(v, wv);
                if (pe & 0x8000u) {
                    acc.sts(xs + (size_t)slot * KP);
                    ++slot;
                    acc.zero();
                }
            }
        }
        __syncwarp();

        // ------------------------------ pass C: (tile, column) partial = sum of the column's slots
        {
            float *pout = partial + (size_t)hd.part0 * KP;
            for (uint32_t idx = lane; idx < hd.C * KP; idx += 32) {
                const uint32_t j = idx / KP, k = idx % KP;
                const uint32_t s0 = cslot[j], s1 = cslot[j + 1];
                float a = 0.0f;
                for (uint32_t sidx = s0; sidx < s1; ++sidx) a += xs[(size_t)sidx * KP + k];
                pout[idx] = a;
            }
        }
        __syncwarp();

        if (next >= n_tiles) break;
        if (!pre) {  // the next tile did not fit beside this one: load it now
            if (lane == 0) {
                mbar_expect_tx(&bar[cur_bar ^ 1], nbytes);
                bulk_g2s(ring, blob + dn_off, nbytes, &bar[cur_bar ^ 1]);
            }
            noff = 0;
        }
        cur_off = noff;
        cur_bytes = nbytes;
        cur_bar ^= 1;
        tile = next;
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
#pragma unroll 4
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

static uint32_t fused_ring_bytes(const polee_handle *h) {
    uint32_t ring = std::max<uint32_t>(5120u, al128(h->ft_max_blob));
    if (const char *e = getenv("POLEE_FUSED_RING")) ring = std::max<uint32_t>(al128((uint32_t)atoi(e)), al128(h->ft_max_blob));
    return ring;
}

template <int KP>
int launch_fused_t(polee_handle *h, const float *x, double *g, bool want_lp, double *lp_partial, float *w_out) {
    const uint32_t ring = fused_ring_bytes(h);
    const FwCarve cv = fw_carve(ring, h->ft_max_rows, h->ft_max_E, h->ft_max_C, h->ft_max_slots, KP);
    const bool weighted = h->ft_row_weight != nullptr;
#define FK_LAUNCH(LPF, WF, WW)                                                                                              \
    do {                                                                                                                    \
        auto kern = k12_fused<KP, LPF, WF, WW>;                                                                             \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv.total);             \
        if (e != cudaSuccess) return h->fail(POLEE_ECUDA, std::string("fused kernel smem: ") + cudaGetErrorString(e));      \
        kern<<<h->ft_grid, FW_WARPS * 32, cv.total, h->stream>>>(h->ft_desc, h->ft_tiles, h->ft_blob, x, h->ft_partial,     \
                                                                 h->ft_row_weight, h->ft_row_of_pos, lp_partial, w_out,     \
                                                                 ring, h->ft_max_rows, h->ft_max_E, h->ft_max_C,            \
                                                                 h->ft_max_slots);                                          \
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

// CTAs (of FW_WARPS warps) for the persistent fused kernel: as many as shared memory lets an SM hold
int fused_grid(polee_handle *h, int KP) {
    const FwCarve cv = fw_carve(fused_ring_bytes(h), h->ft_max_rows, h->ft_max_E, h->ft_max_C, h->ft_max_slots, KP);
    int per_sm = (int)std::max<uint32_t>(1, std::min<uint32_t>(5, (227u * 1024u) / (cv.total + 1024u)));
    if (const char *e = getenv("POLEE_FUSED_CTAS")) per_sm = std::max(1, atoi(e));
    const int ctas_needed = (h->ft_tiles + FW_WARPS - 1) / FW_WARPS;
    return std::max(1, std::min(ctas_needed, h->num_sms * per_sm));
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
