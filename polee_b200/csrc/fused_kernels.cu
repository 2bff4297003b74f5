// fused_kernels.cu -- K1 and K2 of one ADAM step in ONE pass over the matrix ("row tiles", layout in common.cuh).
//
// The split kernels (sparse_kernels.cu) stream the matrix twice per step and round-trip w = 1/p through HBM:
// 2 x nnz x 8 B + 2 x K x m x 4 B = 4.0 GB at C3.  Here a CTA takes a tile of <= 512 consecutive rows (~1.5 k entries,
// one bulk copy), stages x of the tile's <= 255 distinct columns in shared memory, computes p and w for the rows
// (pAt_mul_B!, src/sparse.jl:6-21), keeps w in shared memory, and immediately forms the tile's contribution to
// g = X^T w (pAt_mulinv_B!, src/sparse.jl:25-40) from a second, column-major copy of the entries.  What leaves the SM
// is one partial sum per column segment -- rows arrive sorted by genomic position (src/rnaseq_sample.jl:399-419), so
// a tile touches a handful of columns.  A second, small pass adds the partials of every column (contiguous in the
// column-ordered partial array).  ~1.6 GB per step at C3 instead of 4.0 GB -- but the pass is bound by the
// shared-memory pipe (a 32-byte read of x and one of w per entry), not by HBM, and only beats the split pair where
// columns are short; matrix_setup.cu picks the layout ("K12").
//
// Arithmetic: p = Float32 batches of four products (FFMA2) summed in Float32, w = rcp.approx(p) (the split fast path
// does the same for rows of <= 4 entries and keeps a Float64 row sum above that); column sums: Float32 over a segment
// of <= FT_SEG entries (lane partials + shuffle reduce-scatter), Float64 across segments and tiles.  No atomics, fixed
// orders: run-to-run identical.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "device_utils.cuh"

namespace polee {

namespace {

constexpr int FC_WARPS = 8;
constexpr int FC_THREADS = FC_WARPS * 32;

__host__ __device__ inline uint32_t al128(uint32_t x) { return (x + 127u) & ~127u; }

struct FcCarve {
    uint32_t ring, w_tile, x_loc, total;  // byte offsets; [0, ring) = control block
};
__host__ __device__ inline FcCarve fc_carve(uint32_t ring_bytes, uint32_t max_rows, uint32_t max_C, int KP) {
    FcCarve c;
    c.ring = 128;
    c.w_tile = c.ring + al128(ring_bytes);
    c.x_loc = c.w_tile + al128(((max_rows + 31u) & ~31u) * KP * 4u);
    c.total = c.x_loc + al128(((max_C + 7u) & ~7u) * KP * 4u);
    return c;
}

struct FcCtl {  // shared control block
    uint64_t bar[2];
};

// x of the tile's columns and w of its rows live in shared memory as planes of four draws: item i, draws 4p..4p+3 at
// base + p * stride + i * 4.  A 16-byte access per lane then conflicts only when two lanes of a quarter warp hit items
// that are equal modulo 8 (with 32-byte items it was modulo 4, and both halves).
template <int KP>
__device__ __forceinline__ void sts_planes(float *base, uint32_t stride, uint32_t i, const float *v) {
    if constexpr (KP >= 4) {
#pragma unroll
        for (int q = 0; q < KP / 4; ++q)
            *reinterpret_cast<float4 *>(base + (size_t)q * stride + (size_t)i * 4) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) base[(size_t)i * KP + k] = v[k];
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
    __device__ __forceinline__ void lds(const float *base, uint32_t stride, uint32_t i) {  // plane layout, see sts_planes
        if constexpr (KP >= 4) {
#pragma unroll
            for (int q = 0; q < KP / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(base + (size_t)q * stride + (size_t)i * 4);
                v[2 * q] = make_float2(t.x, t.y);
                v[2 * q + 1] = make_float2(t.z, t.w);
            }
        } else if constexpr (KP == 2) {
            v[0] = *reinterpret_cast<const float2 *>(base + (size_t)i * 2);
        } else {
            v[0] = make_float2(base[i], 0.0f);
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
    __device__ __forceinline__ void add(const Acc &o) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            v[i].x += o.v[i].x;
            v[i].y += o.v[i].y;
        }
    }
};

// Sum acc[0..KP) over the 32 lanes and store the KP totals to out[0..KP).  KP == 8: reduce-scatter (9 shuffles
// instead of 40); other KP: plain butterflies, lane 0 stores.  Fixed order => deterministic.
template <int KP>
__device__ __forceinline__ void warp_reduce_store(float (&acc)[KP], int lane, float *__restrict__ out) {
    if constexpr (KP == 8) {
        float a4[4], a2[2], a1;
        const bool h16 = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = h16 ? acc[i] : acc[i + 4];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            a4[i] = (h16 ? acc[i + 4] : acc[i]) + recv;
        }
        const bool h8 = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = h8 ? a4[i] : a4[i + 2];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
            a2[i] = (h8 ? a4[i + 2] : a4[i]) + recv;
        }
        const bool h4 = lane & 4;
        {
            const float send = h4 ? a2[0] : a2[1];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            a1 = (h4 ? a2[1] : a2[0]) + recv;
        }
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        if ((lane & 3) == 0) out[k] = a1;
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[k] = v;
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KP; ++k) out[k] = acc[k];
        }
    }
}

template <int KP, bool LP, bool WEIGHTED, bool WRITE_W>
__global__ void __launch_bounds__(FC_THREADS, 3)
    k12_fused(const FusedTileDesc *__restrict__ desc, int n_tiles, const unsigned char *__restrict__ blob,
              const float *__restrict__ xf, float *__restrict__ partial, const float *__restrict__ row_weight,
              const uint32_t *__restrict__ row_of_pos, double *__restrict__ lp_partial, float *__restrict__ w_out,
              uint32_t ring_bytes, uint32_t max_rows, uint32_t max_C) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const FcCarve cv = fc_carve(ring_bytes, max_rows, max_C, KP);
    FcCtl &ctl = *reinterpret_cast<FcCtl *>(smraw);
    unsigned char *ring = smraw + cv.ring;
    float *w_tile = reinterpret_cast<float *>(smraw + cv.w_tile);
    float *x_loc = reinterpret_cast<float *>(smraw + cv.x_loc);
    const uint32_t wstride = ((max_rows + 31u) & ~31u) * 4u, xstride = ((max_C + 7u) & ~7u) * 4u;  // floats per plane
    __shared__ double lpsm[FC_WARPS][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tid = threadIdx.x;
    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    if (tid == 0) {
        mbar_init(&ctl.bar[0], 1);
        mbar_init(&ctl.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ring state, computed redundantly (and identically) by every thread; thread 0 issues the copies.
    // The current tile lives at [cur_off, cur_off + cur_bytes).
    const int stride = (int)gridDim.x;
    FusedTileDesc dcur = desc[tile];
    FusedTileDesc dn = tile + stride < n_tiles ? desc[tile + stride] : FusedTileDesc{0, 0, 0};
    uint32_t cur_off = 0, cur_bytes = dcur.bytes, phase = 0;
    uint64_t cur_goff = dcur.off;
    bool cur_issued = false;
    int cur_bar = 0;

    for (;;) {
        __syncthreads();  // everyone is done with the previous tile (its ring space may be reused from here on)
        if (!cur_issued) {
            cur_off = 0;
            if (tid == 0) {
                mbar_expect_tx(&ctl.bar[cur_bar], cur_bytes);
                bulk_g2s(ring, blob + cur_goff, cur_bytes, &ctl.bar[cur_bar]);
            }
        }
        const int next = tile + stride;
        bool pre = false;
        uint32_t noff = 0;
        const uint32_t nbytes = dn.bytes;
        const uint64_t ngoff = dn.off;
        if (next < n_tiles) {  // prefetch the next tile beside the current one when it fits
            const uint32_t cand = cur_off + cur_bytes;
            if (cand + nbytes <= ring_bytes) {
                noff = cand;
                pre = true;
            } else if (nbytes <= cur_off) {
                noff = 0;
                pre = true;
            }
            if (pre && tid == 0) {
                mbar_expect_tx(&ctl.bar[cur_bar ^ 1], nbytes);
                bulk_g2s(ring + noff, blob + ngoff, nbytes, &ctl.bar[cur_bar ^ 1]);
            }
            if (next + stride < n_tiles) dn = desc[next + stride];
        }
        mbar_wait(&ctl.bar[cur_bar], (phase >> cur_bar) & 1u);
        phase ^= 1u << cur_bar;
        const uint32_t my_off = cur_off;
        const unsigned char *b = ring + my_off;
        const FusedHdr hd = *reinterpret_cast<const FusedHdr *>(b);
        const BlobLayout L = blob_layout(hd);
        const uint32_t *cols = reinterpret_cast<const uint32_t *>(b + L.cols);
        const uint32_t *ginfo = reinterpret_cast<const uint32_t *>(b + L.ginfo);
        const uint32_t *dest = reinterpret_cast<const uint32_t *>(b + L.dest);
        const float *valA = reinterpret_cast<const float *>(b + L.valA);
        const float *valB = reinterpret_cast<const float *>(b + L.valB);
        const uint16_t *lrowB = reinterpret_cast<const uint16_t *>(b + L.lrowB);
        const uint16_t *segptr = reinterpret_cast<const uint16_t *>(b + L.segptr);
        const uint8_t *lcolA = b + L.lcolA;

        // ------------------------------ x of the tile's columns -> shared memory (the only gather of the tile)
        for (uint32_t j = tid; j < hd.C; j += FC_THREADS) {
            float xv[KP];
            Vec<KP>::ld(xf + (size_t)cols[j] * KP, xv);
            sts_planes<KP>(x_loc, xstride, j, xv);
        }
        __syncthreads();

        // ------------------------------ pass A: p and w of the tile's rows.  One warp = one group of 32 rows (a dense
        // glen x 32 slab, lane = row); groups are dealt to the warps boustrophedon so that long and short groups mix.
        double lpv[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) lpv[k] = 0.0;
        for (uint32_t round = 0; round * FC_WARPS < hd.G; ++round) {
            const uint32_t g = round * FC_WARPS + ((round & 1u) ? (uint32_t)(FC_WARPS - 1 - warp) : (uint32_t)warp);
            if (g >= hd.G) continue;
            const uint32_t gi = ginfo[g], base = (gi & 0xffffu) + lane, glen = gi >> 16;
            const uint32_t r = g * 32u + lane;
            Acc<KP> fsum;  // Float32 batches of four products, batch sums added in Float32
            fsum.zero();
            uint32_t t = 0;
            for (; t + 4 <= glen; t += 4) {
                Acc<KP> bat;
                bat.zero();
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float v = valA[base + (t + u) * 32u];
                    Acc<KP> xv;
                    xv.lds(x_loc, xstride, lcolA[base + (t + u) * 32u]);
                    bat.fma(v, xv);
                }
                fsum.add(bat);
            }
            if (t < glen) {
                Acc<KP> bat;
                bat.zero();
                for (; t < glen; ++t) {
                    const float v = valA[base + t * 32u];
                    Acc<KP> xv;
                    xv.lds(x_loc, xstride, lcolA[base + t * 32u]);
                    bat.fma(v, xv);
                }
                fsum.add(bat);
            }
            const bool rowok = r < hd.rows;
            float wt = 1.0f;
            if (WEIGHTED && rowok) wt = row_weight[hd.row0 + r];
            float wv[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float p = fsum.get(k);
                const float rc = rcp_approx(p);
                wv[k] = WEIGHTED ? rc * wt : rc;
                if (LP && rowok) lpv[k] += WEIGHTED ? log((double)p) * (double)wt : log((double)p);
            }
            sts_planes<KP>(w_tile, wstride, r, wv);
            if (WRITE_W && rowok) Vec<KP>::st(w_out + (size_t)row_of_pos[hd.row0 + r] * KP, wv);
        }
        if (LP) {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                double v = lpv[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) lpsm[warp][k] = v;
            }
        }
        __syncthreads();
        if (LP && tid < KP) {
            double sacc = 0.0;
            for (int wi = 0; wi < FC_WARPS; ++wi) sacc += lpsm[wi][tid];
            lp_partial[(size_t)tile * KP + tid] = sacc;
        }

        // ------------------------------ pass B: one warp sums one column segment (<= FT_SEG entries) -> one partial
        for (uint32_t sg = warp; sg < hd.S; sg += FC_WARPS) {
            const uint32_t q0 = segptr[sg], q1 = segptr[sg + 1];
            Acc<KP> acc;
            acc.zero();
            for (uint32_t qb = q0; qb < q1; qb += 32u) {  // <= FT_SEG / 32 trips, warp-uniform
                const uint32_t q = qb + lane;
                if (q < q1) {
                    const float v = valB[q];
                    Acc<KP> wv;
                    wv.lds(w_tile, wstride, lrowB[q]);
                    acc.fma(v, wv);
                }
            }
            float a[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) a[k] = acc.get(k);
            warp_reduce_store<KP>(a, lane, partial + (size_t)dest[sg] * KP);
        }

        if (next >= n_tiles) break;
        tile = next;
        cur_off = noff;
        cur_bytes = nbytes;
        cur_goff = ngoff;
        cur_issued = pre;
        cur_bar ^= 1;
    }
}

// g[col] (or a level-2 slot) = sum of <= FT_UNIT partials of one column (contiguous in the column-ordered partial
// array), in tile order.  KP lanes per unit (lane = draw), 32 / KP units per warp: most columns own a handful of
// partials, so a whole warp per unit would mostly wait on its own three dependent loads.
template <int KP>
__global__ void __launch_bounds__(256)
    k_fused_combine1(const FusedUnit *__restrict__ units, int n_units, const float *__restrict__ partial,
                     double *__restrict__ g, double *__restrict__ lvl2) {
    constexpr int G = 32 / KP;
    const int warp_global = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int unit = warp_global * G + lane / KP, k = lane % KP;
    if (unit >= n_units) return;
    const FusedUnit u = units[unit];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four loads in flight; added in a fixed order
    uint32_t i = u.begin;
    for (; i + 4 <= u.end; i += 4) {
        a0 += (double)partial[(size_t)i * KP + k];
        a1 += (double)partial[(size_t)(i + 1) * KP + k];
        a2 += (double)partial[(size_t)(i + 2) * KP + k];
        a3 += (double)partial[(size_t)(i + 3) * KP + k];
    }
    for (; i < u.end; ++i) a0 += (double)partial[(size_t)i * KP + k];
    const double a = (a0 + a1) + (a2 + a3);
    if (u.out < 0)
        g[(size_t)u.col * KP + k] = a;
    else
        lvl2[(size_t)u.out * KP + k] = a;
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
    const uint32_t mb = al128(h->ft_max_blob);
    uint32_t ring = std::max<uint32_t>(mb, std::min<uint32_t>(2 * mb, 48u * 1024u));
    if (const char *e = getenv("POLEE_FUSED_RING")) ring = std::max<uint32_t>(al128((uint32_t)atoi(e)), mb);
    return ring;
}

template <int KP>
int launch_fused_t(polee_handle *h, const float *x, double *g, bool want_lp, double *lp_partial, float *w_out) {
    const uint32_t ring = fused_ring_bytes(h);
    const FcCarve cv = fc_carve(ring, h->ft_max_rows, h->ft_max_C, KP);
    const bool weighted = h->ft_row_weight != nullptr;
#define FK_LAUNCH(LPF, WF, WW)                                                                                           \
    do {                                                                                                                 \
        auto kern = k12_fused<KP, LPF, WF, WW>;                                                                          \
        cudaError_t e = allow_max_smem(kern);                                                                            \
        if (e != cudaSuccess) return h->fail(POLEE_ECUDA, std::string("fused kernel smem: ") + cudaGetErrorString(e));   \
        kern<<<h->ft_grid, FC_THREADS, cv.total, h->stream>>>(h->ft_desc, h->ft_tiles, h->ft_blob, x, h->ft_partial,     \
                                                              h->ft_row_weight, h->ft_row_of_pos, lp_partial, w_out,     \
                                                              ring, h->ft_max_rows, h->ft_max_C);                        \
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
        const int units_per_block = 8 * (32 / KP);
        const int blocks = (h->ft_nunits + units_per_block - 1) / units_per_block;
        k_fused_combine1<KP><<<blocks, 256, 0, h->stream>>>(h->ft_units, h->ft_nunits, h->ft_partial, g, h->ft_lvl2);
    }
    if (h->ft_nmulti > 0) {
        const int blocks = (h->ft_nmulti + 7) / 8;
        k_fused_combine2<KP><<<blocks, 256, 0, h->stream>>>(h->ft_multi, h->ft_nmulti, h->ft_lvl2, g);
    }
    return POLEE_OK;
}

}  // namespace

// CTAs for the persistent fused kernel: what registers and shared memory let an SM hold
int fused_grid(polee_handle *h, int KP) {
    const FcCarve cv = fc_carve(fused_ring_bytes(h), h->ft_max_rows, h->ft_max_C, KP);
    int per_sm = 1;
    auto probe = [&](auto kern) {
        allow_max_smem(kern);
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, FC_THREADS, cv.total) == cudaSuccess && nb > 0) per_sm = nb;
    };
    switch (KP) {
        case 1: probe(k12_fused<1, false, false, false>); break;
        case 2: probe(k12_fused<2, false, false, false>); break;
        case 4: probe(k12_fused<4, false, false, false>); break;
        case 8: probe(k12_fused<8, false, false, false>); break;
        default: probe(k12_fused<16, false, false, false>); break;
    }
    if (const char *e = getenv("POLEE_FUSED_CTAS")) per_sm = std::max(1, atoi(e));
    return std::max(1, std::min(h->ft_tiles, h->num_sms * per_sm));
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
