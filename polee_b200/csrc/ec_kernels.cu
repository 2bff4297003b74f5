// ec_kernels.cu -- the likelihood pass of one ADAM step on the equivalence-class layout (common.cuh, "ec"):
// pAt_mul_B! (src/sparse.jl:6-21), 1 ./ frag_probs, log! + sum (src/likelihood.jl:21-25,43-51) and pAt_mulinv_B!
// (src/sparse.jl:25-40) in ONE pass over the matrix, for the K Monte-Carlo draws of the step at once.
//
// A task = <= ec_nbt(L) blocks of 32 rows of one class (L transcripts).  Per block the two sparse products are the
// dense contractions
//     p[32 x K] = V[32 x L] . x[L x K]            then w = 1 / p            (Float64)
//     g[L x K] += V^T[L x 32] . w[32 x K]                                   (Float64, accumulated over the task)
// done with the FP64 tensor-core MMA (mma.sync m8n8k4 f64 -> SASS DMMA).  The MMA is used for its operand layout, not
// for its peak (DMMA and DFMA both measure 37 TFLOP/s on B200, tools/ubench/fp64_rates.cu): the accumulator fragment
// spreads the L x K column sums over the lanes, so a task needs L/4 registers of accumulators instead of L x K per lane
// and NO cross-lane reduction, and each MMA consumes one conflict-free 128-byte shared-memory read of values.
// Products are exact (Float32 x Float32 in Float64), sums are Float64 throughout -- at least the reference's
// "Float32 product, Float64 accumulation"  -- and 1/p is MUFU.RCP64H + one Newton step (8.5e-13).
//
// Every warp is its own pipeline: it owns a ring of EC_STAGES shared-memory stages fed by 1-D bulk copies
// (cp.async.bulk -> UBLKCP) completing on mbarriers, takes tasks gw, gw + G, gw + 2G, ... (static, so results do not
// depend on scheduling), and never synchronises with another warp.  Output: one Float64 partial per (task, column) at
// its slot of the column-ordered partial array; k_ec_combine1/2 add each column's partials in a fixed order.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "device_utils.cuh"

namespace polee {

namespace {

constexpr int EC_WARPS = 8;
constexpr uint32_t EC_STAGE_STRIDE = (EC_STAGE_BYTES + 127u) & ~127u;

template <int KP>
struct EcCfg {
    static constexpr int NT = KP > 8 ? KP / 8 : 1;          // 8-draw tiles (the MMA's N)
    static constexpr int STAGES = KP > 8 ? 2 : 3;
    static constexpr uint32_t W_BYTES = 32u * 8u * NT * 8u;  // w of one block: [32 rows][8 NT] Float64
    static constexpr uint32_t WARP_BYTES = STAGES * EC_STAGE_STRIDE + W_BYTES;
    static constexpr uint32_t SMEM = EC_WARPS * WARP_BYTES + 256u;  // + the mbarriers
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// 1 / p: MUFU.RCP64H seed (9e-7) + one Newton step (8.5e-13 measured).  p = 0 / inf / NaN keep the seed's inf / 0 / NaN.
__device__ __forceinline__ double rcp64(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);
    return (e == e) ? fma(r, e, r) : r;
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// One task.  MAXLC = compile-time bound on the task's 4-column chunks (L <= 4 * MAXLC); loops are fully unrolled and
// guarded by the (warp-uniform) real counts so that every fragment lives in a register.
template <int KP, int MAXLC, bool LP, bool WEIGHTED, bool WRITE_W>
__device__ __forceinline__ void ec_task(const unsigned char *__restrict__ stage, double *__restrict__ w_sm, int task,
                                        const float *__restrict__ xf, double *__restrict__ partial,
                                        const float *__restrict__ slot_weight, const uint32_t *__restrict__ row_of_slot,
                                        double *__restrict__ lp_partial, float *__restrict__ w_out, int lane) {
    constexpr int NT = EcCfg<KP>::NT;
    constexpr int MAXLT = (MAXLC + 1) / 2;
    constexpr int WROW = 8 * NT;  // doubles per row of w_sm
    const EcHdr hd = *reinterpret_cast<const EcHdr *>(stage);
    const uint32_t L = hd.L, Lp = ec_lp(L), nlc = Lp >> 2, nlt = (nlc + 1u) >> 1;
    const uint32_t *cols = reinterpret_cast<const uint32_t *>(stage + 16);
    const uint32_t *dest = cols + Lp;
    const float *V = reinterpret_cast<const float *>(stage + 16 + 8u * Lp);
    const int g = lane >> 2, t = lane & 3;

    // x of the task's columns as B fragments: (k-row = column 4 lc + t, n = draw g)
    double xb[MAXLC][NT];
#pragma unroll
    for (int lc = 0; lc < MAXLC; ++lc) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) xb[lc][nt] = 0.0;
        if ((uint32_t)lc < nlc) {
            const uint32_t col = cols[lc * 4 + t];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                if (nt * 8 + g < KP) xb[lc][nt] = (double)__ldg(xf + (size_t)col * KP + nt * 8 + g);
        }
    }
    constexpr int CH = MAXLT >= 4 ? 1 : 2;  // independent accumulator chains per tile (hide the MMA latency)
    double d[MAXLT][NT][CH][2];  // [column tile][draw tile][chain][fragment]
#pragma unroll
    for (int lt = 0; lt < MAXLT; ++lt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c2 = 0; c2 < CH; ++c2) d[lt][nt][c2][0] = d[lt][nt][c2][1] = 0.0;
    double lpv[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) lpv[nt][0] = lpv[nt][1] = 0.0;

    // position of this lane's A-fragment element inside a 32-float chunk
    const uint32_t posA = (uint32_t)lane;                                             // p pass: (row g, column t)
    const uint32_t posG0 = (uint32_t)(((t * 4) + (g & 3)) ^ ((g >> 2) << 4));          // g pass, even row chunk
    const uint32_t posG1 = (uint32_t)((((4 + t) * 4) + (g & 3)) ^ ((g >> 2) << 4));    // g pass, odd row chunk

    for (uint32_t b = 0; b < hd.nb; ++b) {
        const float *Vb = V + (size_t)b * Lp * 32u;
        // ---------------- p = V x  (rows 8 mt + g, draws 2 t, 2 t + 1)
        double c[4][NT][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;
#pragma unroll
        for (int lc = 0; lc < MAXLC; ++lc) {
            if ((uint32_t)lc < nlc) {
                const float *ch = Vb + lc * 128 + (posA ^ ((lc & 1) << 4));
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const double a = (double)ch[mt * 32];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(c[mt][nt], a, xb[lc][nt]);
                }
            }
        }
        // ---------------- w = 1 / p (x ks), log p
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const uint32_t row = mt * 8 + g;
            const bool valid = b * 32u + row < hd.rows;
            double wt = 1.0;
            if (WEIGHTED) wt = valid ? (double)slot_weight[hd.slot0 + b * 32u + row] : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double w0 = valid ? rcp64(c[mt][nt][0]) : 0.0;
                double w1 = valid ? rcp64(c[mt][nt][1]) : 0.0;
                if (WEIGHTED) {
                    w0 *= wt;
                    w1 *= wt;
                }
                if (LP && valid) {
                    if (nt * 8 + 2 * t < KP) lpv[nt][0] += WEIGHTED ? wt * log(c[mt][nt][0]) : log(c[mt][nt][0]);
                    if (nt * 8 + 2 * t + 1 < KP) lpv[nt][1] += WEIGHTED ? wt * log(c[mt][nt][1]) : log(c[mt][nt][1]);
                }
                *reinterpret_cast<double2 *>(w_sm + row * WROW + nt * 8 + 2 * t) = make_double2(w0, w1);
                if (WRITE_W && valid) {
                    const size_t r0 = (size_t)row_of_slot[hd.slot0 + b * 32u + row] * KP;
                    if (nt * 8 + 2 * t < KP) w_out[r0 + nt * 8 + 2 * t] = (float)w0;
                    if (nt * 8 + 2 * t + 1 < KP) w_out[r0 + nt * 8 + 2 * t + 1] = (float)w1;
                }
            }
        }
        __syncwarp();
        // ---------------- g += V^T w  (columns 8 lt + g, draws 2 t, 2 t + 1)
        double wb[8][NT];
#pragma unroll
        for (int rc = 0; rc < 8; ++rc)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) wb[rc][nt] = w_sm[(rc * 4 + t) * WROW + nt * 8 + g];
#pragma unroll
        for (int lt = 0; lt < MAXLT; ++lt) {
            if ((uint32_t)lt < nlt) {
                const uint32_t lc = 2u * lt + (uint32_t)(g >> 2);
                const bool has = lc < nlc;
                const float *ch = Vb + (has ? lc : 0u) * 128u;
#pragma unroll
                for (int rc = 0; rc < 8; ++rc) {
                    const float av = ch[(rc >> 1) * 32 + ((rc & 1) ? posG1 : posG0)];
                    const double a = has ? (double)av : 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(d[lt][nt][rc & (CH - 1)], a, wb[rc][nt]);
                }
            }
        }
        __syncwarp();
    }

    // ---------------- the task's column partials -> their slots of the column-ordered partial array
#pragma unroll
    for (int lt = 0; lt < MAXLT; ++lt) {
        const uint32_t l = lt * 8 + g;
        if ((uint32_t)lt < nlt && l < L) {
            double *out = partial + (size_t)dest[l] * KP;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double v0 = CH == 2 ? d[lt][nt][0][0] + d[lt][nt][CH - 1][0] : d[lt][nt][0][0];
                const double v1 = CH == 2 ? d[lt][nt][0][1] + d[lt][nt][CH - 1][1] : d[lt][nt][0][1];
                if constexpr (KP >= 2) {
                    if (nt * 8 + 2 * t < KP) *reinterpret_cast<double2 *>(out + nt * 8 + 2 * t) = make_double2(v0, v1);
                } else {
                    if (t == 0) out[0] = v0;
                }
            }
        }
    }
    if (LP) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double v = lpv[nt][i];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0 && nt * 8 + 2 * t + i < KP) lp_partial[(size_t)task * KP + nt * 8 + 2 * t + i] = v;
            }
    }
}

template <int KP, bool LP, bool WEIGHTED, bool WRITE_W>
__global__ void __launch_bounds__(EC_WARPS * 32, 1)
    k_ec_lik(const EcTaskDesc *__restrict__ desc, int n_tasks, const unsigned char *__restrict__ blob,
             const float *__restrict__ xf, double *__restrict__ partial, const float *__restrict__ slot_weight,
             const uint32_t *__restrict__ row_of_slot, double *__restrict__ lp_partial, float *__restrict__ w_out) {
    using Cfg = EcCfg<KP>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *mine = smraw + 256 + (size_t)warp * Cfg::WARP_BYTES;
    double *w_sm = reinterpret_cast<double *>(mine + STAGES * EC_STAGE_STRIDE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smraw) + warp * STAGES;
    static_assert(EC_WARPS * STAGES * 8 <= 256, "mbarrier block");
    const int gw = blockIdx.x * EC_WARPS + warp, G = gridDim.x * EC_WARPS;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // prologue: the first STAGES tasks of this warp
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const int tk = gw + s * G;
            if (tk < n_tasks) {
                const EcTaskDesc dd = desc[tk];
                mbar_expect_tx(&bars[s], dd.bytes);
                bulk_g2s(mine + s * EC_STAGE_STRIDE, blob + dd.off, dd.bytes, &bars[s]);
            }
        }
    }
    int s = 0;
    uint32_t phase = 0;
    for (int tk = gw; tk < n_tasks; tk += G) {
        // descriptor of the task that will reuse this stage, fetched while the current one is processed
        const int nxt = tk + STAGES * G;
        EcTaskDesc dn{0, 0, 0};
        if (lane == 0 && nxt < n_tasks) dn = desc[nxt];
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        const unsigned char *stage = mine + s * EC_STAGE_STRIDE;
        const uint32_t L = reinterpret_cast<const EcHdr *>(stage)->L;
        if (L <= 8)
            ec_task<KP, 2, LP, WEIGHTED, WRITE_W>(stage, w_sm, tk, xf, partial, slot_weight, row_of_slot, lp_partial, w_out, lane);
        else if (L <= 16)
            ec_task<KP, 4, LP, WEIGHTED, WRITE_W>(stage, w_sm, tk, xf, partial, slot_weight, row_of_slot, lp_partial, w_out, lane);
        else if (L <= 32)
            ec_task<KP, 8, LP, WEIGHTED, WRITE_W>(stage, w_sm, tk, xf, partial, slot_weight, row_of_slot, lp_partial, w_out, lane);
        else
            ec_task<KP, 16, LP, WEIGHTED, WRITE_W>(stage, w_sm, tk, xf, partial, slot_weight, row_of_slot, lp_partial, w_out, lane);
        __syncwarp();  // every lane is done reading the stage
        if (lane == 0 && nxt < n_tasks) {
            mbar_expect_tx(&bars[s], dn.bytes);
            bulk_g2s(mine + s * EC_STAGE_STRIDE, blob + dn.off, dn.bytes, &bars[s]);
        }
        s = (s + 1 == STAGES) ? 0 : s + 1;
    }
}

// g[col] (or a level-2 slot) = [g[col] +] sum of <= FT_UNIT Float64 partials of one column, in task order
template <int KP>
__global__ void __launch_bounds__(256)
    k_ec_combine1(const FusedUnit *__restrict__ units, int n_units, const double *__restrict__ partial, double *__restrict__ g,
                  double *__restrict__ lvl2, int add_to_g) {
    constexpr int GP = 32 / KP;
    const int warp_global = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int unit = warp_global * GP + lane / KP, k = lane % KP;
    if (unit >= n_units) return;
    const FusedUnit u = units[unit];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    uint32_t i = u.begin;
    for (; i + 4 <= u.end; i += 4) {
        a0 += partial[(size_t)i * KP + k];
        a1 += partial[(size_t)(i + 1) * KP + k];
        a2 += partial[(size_t)(i + 2) * KP + k];
        a3 += partial[(size_t)(i + 3) * KP + k];
    }
    for (; i < u.end; ++i) a0 += partial[(size_t)i * KP + k];
    const double a = (a0 + a1) + (a2 + a3);
    if (u.out < 0) {
        double *o = g + (size_t)u.col * KP + k;
        *o = add_to_g ? *o + a : a;
    } else {
        lvl2[(size_t)u.out * KP + k] = a;
    }
}

template <int KP>
__global__ void __launch_bounds__(256)
    k_ec_combine2(const FusedMulti *__restrict__ multi, int n_multi, const double *__restrict__ lvl2, double *__restrict__ g,
                  int add_to_g) {
    const int idx = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (idx >= n_multi) return;
    const int lane = threadIdx.x & 31;
    constexpr int GP = 32 / KP;
    const int grp = lane / KP, k = lane % KP;
    const FusedMulti mc = multi[idx];
    double a = 0.0;
    for (uint32_t i = grp; i < mc.count; i += GP) a += lvl2[(size_t)(mc.first + i) * KP + k];
#pragma unroll
    for (int o = 16; o >= KP; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (grp == 0) {
        double *o = g + (size_t)mc.col * KP + k;
        *o = add_to_g ? *o + a : a;
    }
}

// single CTA, fixed order: out[k] = [out[k] +] sum_t partial[t][k]
__global__ void __launch_bounds__(1024) k_ec_reduce_lp(const double *__restrict__ partial, int count, int KP, double *__restrict__ out,
                                                       int add) {
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
    if (threadIdx.x < KP) out[threadIdx.x] = add ? out[threadIdx.x] + sm[threadIdx.x] : sm[threadIdx.x];
}

template <int KP>
int launch_ec_t(polee_handle *h, const float *x, double *g, bool add_to_g, bool want_lp, double *lp_out, float *w_out) {
    const bool weighted = h->ec_slot_weight != nullptr;
#define EK_LAUNCH(LPF, WF, WW)                                                                                            \
    do {                                                                                                                  \
        auto kern = k_ec_lik<KP, LPF, WF, WW>;                                                                            \
        cudaError_t e = allow_max_smem(kern);                                                                             \
        if (e != cudaSuccess) return h->fail(POLEE_ECUDA, std::string("ec kernel smem: ") + cudaGetErrorString(e));       \
        kern<<<h->ec_grid, EC_WARPS * 32, EcCfg<KP>::SMEM, h->stream>>>(h->ec_desc, h->ec_tasks, h->ec_blob, x, h->ec_partial, \
                                                                         h->ec_slot_weight, h->ec_row_of_slot,            \
                                                                         h->ec_lp_partial, w_out);                        \
    } while (0)
    if (w_out) {
        if (weighted) EK_LAUNCH(false, true, true); else EK_LAUNCH(false, false, true);
    } else if (want_lp) {
        if (weighted) EK_LAUNCH(true, true, false); else EK_LAUNCH(true, false, false);
    } else {
        if (weighted) EK_LAUNCH(false, true, false); else EK_LAUNCH(false, false, false);
    }
#undef EK_LAUNCH
    if (h->ec_nunits > 0) {
        const int units_per_block = 8 * (32 / KP);
        const int blocks = (h->ec_nunits + units_per_block - 1) / units_per_block;
        k_ec_combine1<KP><<<blocks, 256, 0, h->stream>>>(h->ec_units, h->ec_nunits, h->ec_partial, g, h->ec_lvl2, add_to_g ? 1 : 0);
    }
    if (h->ec_nmulti > 0) {
        const int blocks = (h->ec_nmulti + 7) / 8;
        k_ec_combine2<KP><<<blocks, 256, 0, h->stream>>>(h->ec_multi, h->ec_nmulti, h->ec_lvl2, g, add_to_g ? 1 : 0);
    }
    if (want_lp && lp_out) k_ec_reduce_lp<<<1, 1024, 0, h->stream>>>(h->ec_lp_partial, h->ec_tasks, KP, lp_out, add_to_g ? 1 : 0);
    return POLEE_OK;
}

}  // namespace

int ec_grid(polee_handle *h, int KP) {
    (void)KP;
    int ctas = h->num_sms;  // one CTA of EC_WARPS independent warp pipelines per SM (shared memory bound)
    if (const char *e = getenv("POLEE_EC_CTAS")) ctas = std::max(1, atoi(e));
    return std::max(1, std::min(ctas, (h->ec_tasks + EC_WARPS - 1) / EC_WARPS));
}

int launch_ec(polee_handle *h, const float *x, double *g, bool add_to_g, bool want_lp, double *lp_out, float *w_out, int KP) {
    switch (KP) {
        case 1: return launch_ec_t<1>(h, x, g, add_to_g, want_lp, lp_out, w_out);
        case 2: return launch_ec_t<2>(h, x, g, add_to_g, want_lp, lp_out, w_out);
        case 4: return launch_ec_t<4>(h, x, g, add_to_g, want_lp, lp_out, w_out);
        case 8: return launch_ec_t<8>(h, x, g, add_to_g, want_lp, lp_out, w_out);
        case 16: return launch_ec_t<16>(h, x, g, add_to_g, want_lp, lp_out, w_out);
        default: return h->fail(POLEE_EINVAL, "unsupported number of MC draws (1..16)");
    }
}

}  // namespace polee
