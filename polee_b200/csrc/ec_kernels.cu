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
// and NO cross-lane reduction, each MMA consumes one conflict-free 128-byte shared-memory read of values, and with the
// p pass computed transposed the reciprocals are already the g pass's B fragments (w never leaves the registers).
// Products are exact (Float32 x Float32 in Float64), sums are Float64 throughout -- at least the reference's
// "Float32 product, Float64 accumulation"  -- and 1/p is MUFU.RCP64H + one Newton step (8.5e-13).
//
// Every warp is its own pipeline: it owns EC_STAGES shared-memory stages fed by 1-D bulk copies
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

constexpr int EC_WARPS = 11;
constexpr int EC_STAGES = 2;
constexpr uint32_t EC_STAGE_STRIDE = (EC_STAGE_BYTES + 127u) & ~127u;
constexpr uint32_t EC_CTL_BYTES = 384;  // 128 bytes of zeros (the "no such chunk" operand) + the mbarriers

template <int KP>
struct EcCfg {
    static constexpr int NT = KP > 8 ? KP / 8 : 1;  // 8-draw tiles (the MMA's M in the p pass, N in the g pass)
    static constexpr uint32_t WARP_BYTES = EC_STAGES * EC_STAGE_STRIDE;
    static constexpr uint32_t SMEM = EC_WARPS * WARP_BYTES + EC_CTL_BYTES;
};
static_assert(EC_WARPS * EC_STAGES * 8 <= EC_CTL_BYTES - 128, "mbarrier block");

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// 1 / p: MUFU.RCP64H seed (9e-7) + one Newton step (8.5e-13 measured, tools/ubench/fp64_rates.cu).
// p = 0 gives NaN where the reference's 1/p gives Inf: non-finite either way (-> POLEE_ENONFINITE).
__device__ __forceinline__ double rcp64(double p) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double e = fma(-p, r, 1.0);
    return fma(r, e, r);
}

// Float32 -> Float64 with three integer instructions instead of F2F (which shares the 16-lane XU pipe with MUFU):
// exact for normal numbers of either sign; +0 becomes 2^-127 (only the layout's zero padding is 0, and a 1e-39
// addend to a row sum >= 1e-22 is far below Float64 resolution); Inf / NaN become large finite numbers -- the g pass
// converts the same values with F2F, so a non-finite matrix entry still poisons the gradient and is reported.
__device__ __forceinline__ double widen_bits(float f) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t hi = (((u >> 3) & 0x0FFFFFFFu) + 0x38000000u) | (u & 0x80000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}

// One task.  NLC = the task's 4-column chunks; EXACT: nlc == NLC (no guards at all), else nlc <= NLC and the fully
// unrolled loops are guarded by the warp-uniform real count.
//
// p pass, transposed:  p^T[k][r] = sum_l x^T[k][l] V^T[l][r]   (M = 8 draws, N = 8 rows, K = 4 columns per MMA)
//     A (row g, col t) = x[column 4 lc + t][draw g]                    -- registers, loaded once per task
//     B (row t, col g) = V[row 8 j + g][column 4 lc + t]               -- one 128-byte chunk (lc, j), lane order
//     C: thread (g, t) holds p[rows 8 j + 2 t, 8 j + 2 t + 1][draw g]
// g pass:  D[l][k] += sum_r V^T[l][r] w[r][k]                   (M = 8 columns, N = 8 draws, K = 4 rows per MMA)
//     the four K slots of MMA (j, i) are the rows 8 j + 2 t + i, t = 0..3, so that
//     B (row t, col g) = w[row 8 j + 2 t + i][draw g] = 1 / (the thread's own C fragment)  -- no data movement
//     A (row g, col t) = V[row 8 j + 2 t + i][column 8 lt + g]          -- chunk (2 lt + (g >> 2), j)
//     D: thread (g, t) holds g[column 8 lt + g][draws 2 t, 2 t + 1]
template <int KP, int NLC, bool EXACT, bool LP, bool WEIGHTED, bool WRITE_W>
__device__ __forceinline__ void ec_task(const unsigned char *__restrict__ stage, const float *__restrict__ zero_chunk,
                                        int task, const float *__restrict__ xf, double *__restrict__ partial,
                                        const float *__restrict__ slot_weight, const uint32_t *__restrict__ row_of_slot,
                                        double *__restrict__ lp_partial, float *__restrict__ w_out, int lane) {
    constexpr int NT = EcCfg<KP>::NT;
    constexpr int NLT = (NLC + 1) / 2;
    const EcHdr hd = *reinterpret_cast<const EcHdr *>(stage);
    const uint32_t L = hd.L, Lp = ec_lp(L), nlc = EXACT ? (uint32_t)NLC : (Lp >> 2), nlt = (nlc + 1u) >> 1;
    const uint32_t *cols = reinterpret_cast<const uint32_t *>(stage + 16);
    const uint32_t *dest = cols + Lp;
    const float *V = reinterpret_cast<const float *>(stage + 16 + 8u * Lp);
    const int g = lane >> 2, t = lane & 3;

    double xa[NLC][NT];
#pragma unroll
    for (int lc = 0; lc < NLC; ++lc) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) xa[lc][nt] = 0.0;
        if (EXACT || (uint32_t)lc < nlc) {
            const uint32_t col = cols[lc * 4 + t];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                if (nt * 8 + g < KP) xa[lc][nt] = (double)__ldg(xf + (size_t)col * KP + nt * 8 + g);
        }
    }
    double d[NLT][NT][2][2];  // [column tile][draw tile][chain i][fragment]
#pragma unroll
    for (int lt = 0; lt < NLT; ++lt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) d[lt][nt][0][0] = d[lt][nt][0][1] = d[lt][nt][1][0] = d[lt][nt][1][1] = 0.0;
    double lpv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) lpv[nt] = 0.0;

    // g pass: element offsets inside a chunk (floats), and which chunk of a column tile this lane reads
    const uint32_t posG0 = (uint32_t)(((2 * t) * 4 + (g & 3)) ^ ((g >> 2) << 4));
    const uint32_t posG1 = (uint32_t)(((2 * t + 1) * 4 + (g & 3)) ^ ((g >> 2) << 4));

    for (uint32_t b = 0; b < hd.nb; ++b) {
        const float *Vb = V + (size_t)b * Lp * 32u;
        // ---------------- p^T = x^T V^T
        double c[4][NT][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) c[j][nt][0] = c[j][nt][1] = 0.0;
#pragma unroll
        for (int lc = 0; lc < NLC; ++lc) {
            if (EXACT || (uint32_t)lc < nlc) {
                const float *ch = Vb + lc * 128 + ((uint32_t)lane ^ ((lc & 1) << 4));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double v = widen_bits(ch[j * 32]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(c[j][nt], xa[lc][nt], v);
                }
            }
        }
        // ---------------- w = 1 / p (x ks), log p: c becomes w in place
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (LP) {
                        const uint32_t row = b * 32u + 8 * j + 2 * t + i;
                        if (row < hd.rows && nt * 8 + g < KP)
                            lpv[nt] += WEIGHTED ? (double)slot_weight[hd.slot0 + row] * log(c[j][nt][i]) : log(c[j][nt][i]);
                    }
                    c[j][nt][i] = rcp64(c[j][nt][i]);
                }
        if (WEIGHTED) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t row = b * 32u + 8 * j + 2 * t + i;
                    const double wt = row < hd.rows ? (double)slot_weight[hd.slot0 + row] : 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) c[j][nt][i] *= wt;
                }
        }
        if (b * 32u + 32u > hd.rows || KP < 8) {  // padding rows (last block of a class) and padding draws: w = 0
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool rok = b * 32u + 8 * j + 2 * t + i < hd.rows;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if (!rok || nt * 8 + g >= KP) c[j][nt][i] = 0.0;
                }
        }
        if (WRITE_W) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t row = b * 32u + 8 * j + 2 * t + i;
                    if (row < hd.rows) {
                        const size_t r0 = (size_t)row_of_slot[hd.slot0 + row] * KP;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            if (nt * 8 + g < KP) w_out[r0 + nt * 8 + g] = (float)c[j][nt][i];
                    }
                }
        }
        // ---------------- g += V^T w
#pragma unroll
        for (int lt = 0; lt < NLT; ++lt) {
            if (EXACT || (uint32_t)lt < nlt) {
                const uint32_t lc = 2u * lt + (uint32_t)(g >> 2);
                const float *ch = (lc < nlc) ? Vb + lc * 128u : zero_chunk;
                const uint32_t jstep = (lc < nlc) ? 32u : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double a = (double)ch[j * jstep + (i ? posG1 : posG0)];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma(d[lt][nt][i], a, c[j][nt][i]);
                    }
            }
        }
    }

    // ---------------- the task's column partials -> their slots of the column-ordered partial array
#pragma unroll
    for (int lt = 0; lt < NLT; ++lt) {
        const uint32_t l = lt * 8 + g;
        if ((EXACT || (uint32_t)lt < nlt) && l < L) {
            double *out = partial + (size_t)dest[l] * KP;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double v0 = d[lt][nt][0][0] + d[lt][nt][1][0], v1 = d[lt][nt][0][1] + d[lt][nt][1][1];
                if constexpr (KP >= 2) {
                    if (nt * 8 + 2 * t < KP) *reinterpret_cast<double2 *>(out + nt * 8 + 2 * t) = make_double2(v0, v1);
                } else {
                    if (t == 0) out[0] = v0;
                }
            }
        }
    }
    if (LP) {  // lanes (g, t = 0..3) hold partial sums of draw g
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            double v = lpv[nt];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (t == 0 && nt * 8 + g < KP) lp_partial[(size_t)task * KP + nt * 8 + g] = v;
        }
    }
}

template <int KP, bool LP, bool WEIGHTED, bool WRITE_W>
__global__ void __launch_bounds__(EC_WARPS * 32, 1)
    k_ec_lik(const EcTaskDesc *__restrict__ desc, int n_tasks, const unsigned char *__restrict__ blob,
             const float *__restrict__ xf, double *__restrict__ partial, const float *__restrict__ slot_weight,
             const uint32_t *__restrict__ row_of_slot, double *__restrict__ lp_partial, float *__restrict__ w_out) {
    using Cfg = EcCfg<KP>;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *zero_chunk = reinterpret_cast<const float *>(smraw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + 128) + warp * EC_STAGES;
    unsigned char *mine = smraw + EC_CTL_BYTES + (size_t)warp * Cfg::WARP_BYTES;
    const int gw = blockIdx.x * EC_WARPS + warp, G = gridDim.x * EC_WARPS;
    if (threadIdx.x < 32) reinterpret_cast<float *>(smraw)[threadIdx.x] = 0.0f;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < EC_STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only block-wide barrier: zero chunk + mbarriers visible
    if (lane == 0) {  // prologue: the first EC_STAGES tasks of this warp
#pragma unroll
        for (int s = 0; s < EC_STAGES; ++s) {
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
        const int nxt = tk + EC_STAGES * G;
        EcTaskDesc dn{0, 0, 0};
        if (lane == 0 && nxt < n_tasks) dn = desc[nxt];
        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        const unsigned char *stage = mine + s * EC_STAGE_STRIDE;
        const uint32_t nlc = ec_lp(reinterpret_cast<const EcHdr *>(stage)->L) >> 2;
#define EC_CALL(N, EX) ec_task<KP, N, EX, LP, WEIGHTED, WRITE_W>(stage, zero_chunk, tk, xf, partial, slot_weight, row_of_slot, lp_partial, w_out, lane)
        switch (nlc) {
            case 1: EC_CALL(1, true); break;
            case 2: EC_CALL(2, true); break;
            case 3: EC_CALL(3, true); break;
            case 4: EC_CALL(4, true); break;
            default:
                if (nlc <= 8) EC_CALL(8, false); else EC_CALL(16, false);
        }
#undef EC_CALL
        __syncwarp();  // every lane is done reading the stage
        if (lane == 0 && nxt < n_tasks) {
            mbar_expect_tx(&bars[s], dn.bytes);
            bulk_g2s(mine + s * EC_STAGE_STRIDE, blob + dn.off, dn.bytes, &bars[s]);
        }
        s = (s + 1 == EC_STAGES) ? 0 : s + 1;
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
