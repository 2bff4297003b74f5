/*
 * polee_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY
 * (see polee_oracle.h).  Compile with -ffp-contract=off: Julia never fuses a*b+c, so neither
 * may this file.  OpenMP `parallel for schedule(static)` stands where the reference has
 * `Threads.@threads` (static chunking), everything else is serial like the reference.
 */
#include "polee_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1; the reference's launcher uses every core (polee:8-12) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* =============================== sparse.jl =============================== */

/* sparse.jl:6-21.  `y[j] = zero(T)` then `y[j] += x[i] * nzval[k]`: the product of two Float32
 * is a Float32 (rounded), the accumulation happens in y's eltype (Float64). */
void orc_pAt_mul_B_f32(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, const float *x) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < ncols; ++j) {
        double acc = 0.0;
        for (int64_t k = (int64_t)colptr[j] - 1; k < (int64_t)colptr[j + 1] - 1; ++k) {
            int64_t i = (int64_t)rowval[k] - 1;
            float prod = x[i] * nzval[k];
            acc += (double)prod;
        }
        y[j] = acc;
    }
}

/* sparse.jl:6-21 with x::Vector{Float64} (likelihood.jl:82): Float64 * Float32 -> Float64. */
void orc_pAt_mul_B_f64(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, const double *x) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < ncols; ++j) {
        double acc = 0.0;
        for (int64_t k = (int64_t)colptr[j] - 1; k < (int64_t)colptr[j + 1] - 1; ++k) {
            int64_t i = (int64_t)rowval[k] - 1;
            acc += x[i] * (double)nzval[k];
        }
        y[j] = acc;
    }
}

/* sparse.jl:25-40: `y[j] += nzval[k] / x[i]` with x Float64 -> Float64 divide and sum. */
void orc_pAt_mulinv_B(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                      const float *nzval, const double *x) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < ncols; ++j) {
        double acc = 0.0;
        for (int64_t k = (int64_t)colptr[j] - 1; k < (int64_t)colptr[j + 1] - 1; ++k) {
            int64_t i = (int64_t)rowval[k] - 1;
            acc += (double)nzval[k] / x[i];
        }
        y[j] = acc;
    }
}

/* likelihood-approximation.jl:407 `SparseMatrixCSC(transpose(X))`: counting transpose; the entries
 * of each output column come out in ascending row order, as Julia's halfperm produces. */
void orc_transpose_csc(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, uint32_t *t_colptr, uint32_t *t_rowval, float *t_nzval) {
    int64_t nnz = (int64_t)colptr[n] - 1;
    int64_t *cursor = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nnz; ++k) cursor[rowval[k]]++; /* count into slot row (1-based) */
    int64_t run = 0;
    for (int64_t i = 0; i < m; ++i) {
        int64_t c = cursor[i + 1];
        t_colptr[i] = (uint32_t)(run + 1);
        cursor[i + 1] = run;
        run += c;
    }
    t_colptr[m] = (uint32_t)(run + 1);
    for (int64_t j = 0; j < n; ++j) {
        for (int64_t k = (int64_t)colptr[j] - 1; k < (int64_t)colptr[j + 1] - 1; ++k) {
            int64_t i = rowval[k];
            int64_t dst = cursor[i]++;
            t_rowval[dst] = (uint32_t)(j + 1);
            t_nzval[dst] = nzval[k];
        }
    }
    free(cursor);
}

/* =============================== likelihood.jl =============================== */

/* Julia's sum(::Vector{Float64}) is pairwise with a 1024-element serial base case
 * (Base.mapreduce_impl); the @simd reassociation inside the base case is not reproduced. */
static double pairwise_sum(const double *a, int64_t lo, int64_t hi) {
    if (hi - lo <= 1024) {
        double s = 0.0;
        for (int64_t i = lo; i < hi; ++i) s += a[i];
        return s;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    return pairwise_sum(a, lo, mid) + pairwise_sum(a, mid, hi);
}

/* likelihood.jl:36-56 */
double orc_log_likelihood(int64_t m, int64_t n, double *frag_probs, double *log_frag_probs,
                          const uint32_t *colptr, const uint32_t *rowval, const float *nzval,
                          const uint32_t *t_colptr, const uint32_t *t_rowval, const float *t_nzval,
                          const float *xs, double *x_grad, int gradonly) {
    orc_pAt_mul_B_f32(frag_probs, m, t_colptr, t_rowval, t_nzval, xs); /* :43 */
    double lp = 0.0;
    if (!gradonly) { /* :46-51, log! is :21-25 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i) log_frag_probs[i] = log(frag_probs[i]);
        lp = pairwise_sum(log_frag_probs, 0, m);
    }
    orc_pAt_mulinv_B(x_grad, n, colptr, rowval, nzval, frag_probs); /* :53 */
    return lp;
}

/* likelihood.jl:59-85 */
double orc_factored_log_likelihood(int64_t m, int64_t n, double *frag_probs, double *log_frag_probs,
                                   const uint32_t *colptr, const uint32_t *rowval, const float *nzval,
                                   const uint32_t *t_colptr, const uint32_t *t_rowval,
                                   const float *t_nzval, const int64_t *ks, const float *xs,
                                   double *x_grad, int gradonly) {
    orc_pAt_mul_B_f32(frag_probs, m, t_colptr, t_rowval, t_nzval, xs); /* :67 */
    double lp = 0.0;
    if (!gradonly) { /* :70-76 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i) log_frag_probs[i] = log(frag_probs[i]) * (double)ks[i];
        lp = pairwise_sum(log_frag_probs, 0, m);
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; ++i) frag_probs[i] = (double)ks[i] / frag_probs[i]; /* :78-80 */
    orc_pAt_mul_B_f64(x_grad, n, colptr, rowval, nzval, frag_probs);               /* :82 */
    return lp;
}

/* likelihood.jl:93-110.  xls Float32, x_scaled_sum Float64, `n * (1/efflens[i])` is Float32
 * (Int * Float32), then divided by the Float64 sum. */
double orc_effective_length_jacobian_adjustment(int64_t n, const float *efflens, const float *xs,
                                                float *xls, double *x_grad) {
    double x_scaled_sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        xls[i] = xs[i] / efflens[i];
        x_scaled_sum += (double)xls[i];
    }
    for (int64_t i = 0; i < n; ++i) xls[i] = (float)((double)xls[i] / x_scaled_sum);
    for (int64_t i = 0; i < n; ++i) {
        float a = (float)n * (1.0f / efflens[i]);
        x_grad[i] -= (double)a / x_scaled_sum;
    }
    return 0.0;
}

/* likelihood.jl:114-159.  `values(gene_transcripts)` iterates a Julia Dict (order unpinned); it only decides the
 * order of the Float64 sum `offdiag_contrib`, which here runs over transcripts 1..n as the reference's loop :144 does. */
double orc_gene_noninformative_prior(int64_t n, const float *efflens, const float *xls, double *xl_grad,
                                     const float *xs, double *x_grad, const orc_genes *genes) {
    for (int64_t i = 0; i < n; ++i) xl_grad[i] = 0.0;                                   /* :118 */
    for (int64_t g = 0; g < genes->num_genes; ++g) {                                     /* :120-133 */
        int64_t b = genes->gene_ptr[g], e = genes->gene_ptr[g + 1], k = e - b;
        if (k > 1) {
            double c = 0.0;
            for (int64_t q = b; q < e; ++q) c += (double)xls[genes->transcripts[q] - 1];
            for (int64_t q = b; q < e; ++q) xl_grad[genes->transcripts[q] - 1] = -(double)(k - 1) / c;
        }
    }
    double x_scaled_sum = 0.0;                                                           /* :137-141 */
    for (int64_t i = 0; i < n; ++i) x_scaled_sum += (double)(xs[i] / efflens[i]);
    double x_scaled_sum_sq = x_scaled_sum * x_scaled_sum;
    double offdiag_contrib = 0.0;                                                        /* :143-147 */
    for (int64_t i = 0; i < n; ++i) offdiag_contrib += -xl_grad[i] * (double)xls[i];
    offdiag_contrib /= x_scaled_sum_sq;
    for (int64_t i = 0; i < n; ++i) {                                                    /* :149-154 */
        float inv = 1.0f / efflens[i];
        double grad_a = xl_grad[i] * ((double)inv / x_scaled_sum);
        double grad_b = (double)inv * offdiag_contrib;
        x_grad[i] += grad_a + grad_b;
    }
    return 0.0;
}

/* =============================== ptt.jl =============================== */

#define IDX(t, r, i) ((t)->index[4 * (size_t)((i)-1) + ((r)-1)]) /* 1-based (row, node) */
#define GRD(t, r, i) ((t)->gradients[2 * (size_t)((i)-1) + ((r)-1)])
#define US(t, i) ((t)->us[(i)-1])

/* ptt.jl:89-116: the first child seen for a parent is its RIGHT child. */
orc_ptt *orc_ptt_new(const int32_t *parent_idxs, const int32_t *js, int64_t num_nodes) {
    orc_ptt *t = (orc_ptt *)malloc(sizeof(orc_ptt));
    t->num_nodes = num_nodes;
    t->index = (int32_t *)calloc((size_t)num_nodes * 4, sizeof(int32_t));
    t->us = (double *)calloc((size_t)num_nodes, sizeof(double));
    t->gradients = (float *)calloc((size_t)num_nodes * 2, sizeof(float));
    for (int64_t i = 1; i <= num_nodes; ++i) {
        IDX(t, 1, i) = js[i - 1];
        int32_t p = parent_idxs[i - 1];
        if (p != 0) {
            if (IDX(t, 3, p) == 0)
                IDX(t, 3, p) = (int32_t)i;
            else
                IDX(t, 2, p) = (int32_t)i;
        }
        IDX(t, 4, i) = p;
    }
    return t;
}

void orc_ptt_free(orc_ptt *t) {
    if (!t) return;
    free(t->index);
    free(t->us);
    free(t->gradients);
    free(t);
}

const int32_t *orc_ptt_index(const orc_ptt *t) { return t->index; }

/* ptt.jl:125-160.  xs is Float32: `xs[o] = us[i]` rounds, `max(xs[o], 1e-16)` promotes to
 * Float64 and the store rounds again. */
double orc_ptt_transform(orc_ptt *t, const double *ys, float *xs, int compute_ladj) {
    double ladj = 0.0;
    US(t, 1) = 1.0;
    int64_t k = 1;
    for (int64_t i = 1; i <= t->num_nodes; ++i) {
        int32_t o = IDX(t, 1, i);
        if (o != 0) {
            float v = (float)US(t, i);
            double w = (double)v;
            xs[o - 1] = (float)(w > 1e-16 ? w : 1e-16);
            continue;
        }
        int32_t l = IDX(t, 2, i), r = IDX(t, 3, i);
        US(t, l) = ys[k - 1] * US(t, i);
        US(t, r) = (1.0 - ys[k - 1]) * US(t, i);
        if (compute_ladj) ladj += log(US(t, i));
        ++k;
    }
    return ladj;
}

/* ptt.jl:167-209.  t.gradients is Float32: (lg + llg) - (rg + rlg) is evaluated in Float32,
 * the rest in Float64 and rounded on store. */
void orc_ptt_transform_gradients(orc_ptt *t, const double *ys, float *y_grad, const double *x_grad) {
    int64_t N = t->num_nodes, n = (N + 1) / 2, k = n - 1;
    for (int64_t i = N; i >= 1; --i) {
        int32_t o = IDX(t, 1, i);
        if (o != 0) {
            GRD(t, 1, i) = (float)x_grad[o - 1];
            GRD(t, 2, i) = 0.0f;
            continue;
        }
        int32_t l = IDX(t, 2, i), r = IDX(t, 3, i);
        float lg = GRD(t, 1, l), llg = GRD(t, 2, l);
        float rg = GRD(t, 1, r), rlg = GRD(t, 2, r);
        float a = lg + llg, b = rg + rlg;
        float d = a - b;
        double y = ys[k - 1];
        y_grad[k - 1] = (float)(US(t, i) * (double)d);
        double g1 = y * (double)lg;
        double g1b = (1.0 - y) * (double)rg;
        GRD(t, 1, i) = (float)(g1 + g1b);
        double g2 = 1.0 / US(t, i);
        double g2a = y * (double)llg;
        double g2b = (1.0 - y) * (double)rlg;
        GRD(t, 2, i) = (float)((g2 + g2a) + g2b);
        --k;
    }
}

/* ptt.jl:217-251, with y_grad::Vector{Float64} (likelihood-approximation.jl:181) */
void orc_ptt_transform_gradients_no_ladj(orc_ptt *t, const double *ys, double *y_grad,
                                         const double *x_grad) {
    int64_t N = t->num_nodes, n = (N + 1) / 2, k = n - 1;
    for (int64_t i = N; i >= 1; --i) {
        int32_t o = IDX(t, 1, i);
        if (o != 0) {
            GRD(t, 1, i) = (float)x_grad[o - 1];
            GRD(t, 2, i) = 0.0f;
            continue;
        }
        int32_t l = IDX(t, 2, i), r = IDX(t, 3, i);
        float lg = GRD(t, 1, l), rg = GRD(t, 1, r);
        float d = lg - rg;
        double y = ys[k - 1];
        y_grad[k - 1] = US(t, i) * (double)d;
        double g1 = y * (double)lg;
        double g1b = (1.0 - y) * (double)rg;
        GRD(t, 1, i) = (float)(g1 + g1b);
        --k;
    }
}

/* ptt.jl:257-285 */
double orc_ptt_inverse_transform(orc_ptt *t, const float *xs, double *ys) {
    int64_t N = t->num_nodes, n = (N + 1) / 2, k = n - 1;
    double ladj = 0.0;
    for (int64_t i = N; i >= 1; --i) {
        int32_t o = IDX(t, 1, i);
        if (o != 0) {
            US(t, i) = (double)xs[o - 1];
            continue;
        }
        int32_t l = IDX(t, 2, i), r = IDX(t, 3, i);
        US(t, i) = US(t, l) + US(t, r);
        ladj -= (double)logf((float)US(t, i));
        ys[k - 1] = US(t, l) / US(t, i);
        --k;
    }
    return ladj;
}

/* ptt.jl:293-309 */
void orc_make_inverse_ptt_params(const int32_t *node_parent_idxs, const int32_t *node_js,
                                 int64_t num_nodes, int32_t *left_index, int32_t *right_index,
                                 int32_t *leaf_index) {
    for (int64_t i = 0; i < num_nodes; ++i) left_index[i] = right_index[i] = -1;
    for (int64_t i = 2; i <= num_nodes; ++i) {
        int32_t p = node_parent_idxs[i - 1];
        if (right_index[p - 1] == -1)
            right_index[p - 1] = (int32_t)(i - 1);
        else
            left_index[p - 1] = (int32_t)(i - 1);
    }
    for (int64_t i = 0; i < num_nodes; ++i) leaf_index[i] = node_js[i] - 1;
}

/* hclust.jl:477-489 (list_nodes) followed by order_nodes :361-389.  The stack build pops a = n,
 * b = n-1, pushes I(a,b); then a = I, b = n-2 ...; the root is I(chain, leaf 1).  DFS emits the
 * right child first: root, leaf 1, I, leaf 2, I, ..., leaf n-1, leaf n. */
void orc_list_nodes(int64_t n, int32_t *parent_idxs, int32_t *js) {
    int64_t N = 2 * n - 1, pos = 0;
    int32_t parent = 0;
    for (int64_t leaf = 1; leaf <= n - 1; ++leaf) {
        parent_idxs[pos] = parent; /* internal node */
        js[pos] = 0;
        int32_t me = (int32_t)(pos + 1);
        ++pos;
        parent_idxs[pos] = me; /* its right child: leaf `leaf` */
        js[pos] = (int32_t)leaf;
        ++pos;
        parent = me;
    }
    parent_idxs[pos] = parent;
    js[pos] = (int32_t)n;
    ++pos;
    (void)N;
}

/* =============================== reparameterisation =============================== */

/* sinh_arcsinh.jl:10-23 (all Float32).  The reference's threaded `ladj +=` is racy
 * (SURVEY App. C3); the serial sum is restated. */
double orc_sinh_asinh_transform(int64_t nm1, const float *alpha, const float *zs0, float *zs,
                                int compute_ladj) {
    double ladj = 0.0; /* `ladj = 0.0f0; ladj += <Float64>` is type-unstable: it is a Float64 after the first add */
#pragma omp parallel for schedule(static) if (!compute_ladj)
    for (int64_t i = 0; i < nm1; ++i) {
        float c = alpha[i] + asinhf(zs0[i]);
        zs[i] = sinhf(c);
        if (compute_ladj) {
            /* log(cosh(c)) Float32 - 0.5 * log1p(z0^2): 0.5 is Float64 -> Float64, `ladj +=` */
            double term = (double)logf(coshf(c)) - 0.5 * (double)log1pf(zs0[i] * zs0[i]);
            ladj += term;
        }
    }
    return ladj;
}

/* logitnormal.jl:2-20: logistic(x) = inv(1 + exp(-x)) in Float32, stored into Float64 ys. */
double orc_logit_normal_transform(int64_t nm1, const float *mu, const float *sigma, const float *zs,
                                  double *ys, int compute_ladj) {
    double ladj = 0.0;
    for (int64_t i = 0; i < nm1; ++i) {
        float prod = zs[i] * sigma[i];
        float x = mu[i] + prod;
        float e = expf(-x);
        float y = 1.0f / (1.0f + e);
        ys[i] = (double)y;
        if (compute_ladj) {
            /* log(sigma[i] * ys[i] * (1 - ys[i])): ys is Float64 here */
            double v = ((double)sigma[i] * ys[i]) * (1.0 - ys[i]);
            ladj += log(v);
        }
    }
    return ladj;
}

/* logitnormal.jl:38-55 (8-argument method); Float32 accumulators, Float64 intermediates. */
void orc_logit_normal_transform_gradients(int64_t nm1, const float *zs, const double *ys,
                                          const float *mu, const float *sigma, const float *y_grad,
                                          float *z_grad, float *mu_grad, float *sigma_grad) {
    (void)mu;
    for (int64_t i = 0; i < nm1; ++i) {
        double y = ys[i];
        double dy_dmu = y * (1.0 - y);
        mu_grad[i] = (float)((double)mu_grad[i] + dy_dmu * (double)y_grad[i]);
        double dy_dsigma = (y * (1.0 - y)) * (double)zs[i];
        sigma_grad[i] = (float)((double)sigma_grad[i] + dy_dsigma * (double)y_grad[i]);
        double dy_dz = (y * (1.0 - y)) * (double)sigma[i];
        z_grad[i] = (float)((double)z_grad[i] + dy_dz * (double)y_grad[i]);
        /* ladj gradients */
        mu_grad[i] = (float)((double)mu_grad[i] + (1.0 - 2.0 * y));
        float inv_sigma = 1.0f / sigma[i];
        double t2 = (double)zs[i] * (1.0 - 2.0 * y);
        sigma_grad[i] = (float)((double)sigma_grad[i] + ((double)inv_sigma + t2));
        z_grad[i] = (float)((double)z_grad[i] + (double)sigma[i] * (1.0 - 2.0 * y));
    }
}

/* sinh_arcsinh.jl:29-38 (all Float32) */
void orc_sinh_asinh_transform_gradients(int64_t nm1, const float *zs0, const float *alpha,
                                        const float *z_grad, float *alpha_grad) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nm1; ++i) {
        float c = alpha[i] + asinhf(zs0[i]);
        float dz_dalpha = coshf(c);
        float p = dz_dalpha * z_grad[i];
        alpha_grad[i] = alpha_grad[i] + p;
        alpha_grad[i] = alpha_grad[i] + tanhf(c);
    }
}

/* =============================== ADAM =============================== */

/* likelihood-approximation.jl:107-110 */
double orc_adam_learning_rate(int64_t step_num) {
    double lr = ORC_ADAM_INITIAL_LEARNING_RATE * exp(-ORC_ADAM_LEARNING_RATE_DECAY * (double)step_num);
    return lr > ORC_ADAM_MIN_LEARNING_RATE ? lr : ORC_ADAM_MIN_LEARNING_RATE;
}

/* likelihood-approximation.jl:116-130; grad::Vector{Float32}: grad^2 is Float32 */
void orc_adam_update_mv(int64_t len, float *ms, float *vs, const float *grad, int64_t step_num) {
    if (step_num == 1) {
        for (int64_t i = 0; i < len; ++i) {
            ms[i] = grad[i];
            vs[i] = grad[i] * grad[i];
        }
    } else {
        for (int64_t i = 0; i < len; ++i) {
            double a = ORC_ADAM_RM * (double)ms[i];
            double b = (1.0 - ORC_ADAM_RM) * (double)grad[i];
            ms[i] = (float)(a + b);
            float g2 = grad[i] * grad[i];
            double c = ORC_ADAM_RV * (double)vs[i];
            double d = (1.0 - ORC_ADAM_RV) * (double)g2;
            vs[i] = (float)(c + d);
        }
    }
}

/* likelihood-approximation.jl:136-146 (gradient ASCENT) */
void orc_adam_update_params(int64_t len, float *params, const float *ms, const float *vs,
                            double learning_rate, int64_t step_num, double max_step_size) {
    double m_denom = 1.0 - pow(ORC_ADAM_RM, (double)step_num);
    double v_denom = 1.0 - pow(ORC_ADAM_RV, (double)step_num);
    for (int64_t i = 0; i < len; ++i) {
        double param_m = (double)ms[i] / m_denom;
        double param_v = (double)vs[i] / v_denom;
        double delta = (learning_rate * param_m) / (sqrt(param_v) + ORC_ADAM_EPS);
        if (delta < -max_step_size) delta = -max_step_size;
        if (delta > max_step_size) delta = max_step_size;
        params[i] = (float)((double)params[i] + delta);
    }
}

/* =============================== noise =============================== */

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* "polee-philox-v1": one Philox4x32-10 call per element, counter = (i, draw, step, 0), key = seed;
 * z = sqrt(-2 ln u1) cos(2 pi u2) from the first two output words (Float32 arithmetic). */
void orc_noise_fill(uint64_t seed, int64_t step, int64_t draw, int64_t nm1, float *zs0) {
    for (int64_t i = 0; i < nm1; ++i) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)draw, (uint32_t)step, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float r = sqrtf(-2.0f * logf(u1));
        zs0[i] = r * cosf(6.28318530717958647692f * u2);
    }
}

/* =============================== fits =============================== */

typedef struct {
    int64_t m, n;
    const uint32_t *colptr, *rowval;
    const float *nzval;
    uint32_t *t_colptr, *t_rowval;
    float *t_nzval;
    double *frag_probs, *log_frag_probs;
} orc_model;

static void model_init(orc_model *M, int64_t m, int64_t n, const uint32_t *colptr,
                       const uint32_t *rowval, const float *nzval) {
    int64_t nnz = (int64_t)colptr[n] - 1;
    M->m = m; M->n = n; M->colptr = colptr; M->rowval = rowval; M->nzval = nzval;
    M->t_colptr = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(m + 1));
    M->t_rowval = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nnz > 0 ? nnz : 1));
    M->t_nzval = (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1));
    M->frag_probs = (double *)calloc((size_t)m, sizeof(double));
    M->log_frag_probs = (double *)calloc((size_t)m, sizeof(double));
    orc_transpose_csc(m, n, colptr, rowval, nzval, M->t_colptr, M->t_rowval, M->t_nzval);
}

static void model_free(orc_model *M) {
    free(M->t_colptr); free(M->t_rowval); free(M->t_nzval);
    free(M->frag_probs); free(M->log_frag_probs);
}

/* initial values, likelihood-approximation.jl:451-456 */
static void lsn_init(orc_ptt *t, int64_t n, float *mu, float *omega, float *alpha) {
    float *xs0 = (float *)malloc(sizeof(float) * (size_t)n);
    double *ys = (double *)malloc(sizeof(double) * (size_t)(n - 1));
    for (int64_t i = 0; i < n; ++i) xs0[i] = 1.0f / (float)n;
    orc_ptt_inverse_transform(t, xs0, ys);
    for (int64_t i = 0; i < n - 1; ++i) {
        mu[i] = (float)log(ys[i] / (1.0 - ys[i])); /* logit in Float64, map! rounds to Float32 */
        omega[i] = logf(0.1f);
        alpha[i] = 0.0f;
    }
    free(xs0); free(ys);
}

/* scratch for one draw */
typedef struct {
    float *zs, *xs, *xls, *y_grad, *z_grad, *sigma_grad, *sigma;
    double *ys, *x_grad, *xl_grad;
} draw_ws;

static void ws_init(draw_ws *w, int64_t n) {
    size_t nm1 = (size_t)(n - 1);
    w->zs = (float *)calloc(nm1, 4); w->xs = (float *)calloc((size_t)n, 4);
    w->xls = (float *)calloc((size_t)n, 4); w->y_grad = (float *)calloc(nm1, 4);
    w->z_grad = (float *)calloc(nm1, 4); w->sigma_grad = (float *)calloc(nm1, 4);
    w->sigma = (float *)calloc(nm1, 4);
    w->ys = (double *)calloc(nm1, 8); w->x_grad = (double *)calloc((size_t)n, 8);
    w->xl_grad = (double *)calloc((size_t)n, 8);
}
static void ws_free(draw_ws *w) {
    free(w->zs); free(w->xs); free(w->xls); free(w->y_grad); free(w->z_grad);
    free(w->sigma_grad); free(w->sigma); free(w->ys); free(w->x_grad); free(w->xl_grad);
}

/* body of the MC loop, likelihood-approximation.jl:512-549 (and :322-355 when ks != NULL).
 * Accumulates into mu_grad / omega_grad / alpha_grad, returns this draw's "elbo". */
static double lsn_draw_body(orc_model *M, orc_ptt *t, const int64_t *ks, const float *efflens,
                            int gradonly, int use_efflen_jacobian, const orc_genes *genes, const float *mu,
                            const float *alpha, const float *zs0, draw_ws *w, float *mu_grad,
                            float *omega_grad, float *alpha_grad) {
    int64_t n = M->n, nm1 = n - 1;
    const double eps = 1e-10;
    memset(w->x_grad, 0, sizeof(double) * (size_t)n);        /* :512-515 */
    memset(w->y_grad, 0, sizeof(float) * (size_t)nm1);
    memset(w->z_grad, 0, sizeof(float) * (size_t)nm1);
    memset(w->sigma_grad, 0, sizeof(float) * (size_t)nm1);
    int ladj_on = ks ? 0 : !gradonly; /* factored variant always passes Val(false) / Val(true) */
    int lik_gradonly = ks ? 1 : gradonly;

    double skew_ladj = orc_sinh_asinh_transform(nm1, alpha, zs0, w->zs, ladj_on);      /* :521 */
    double ln_ladj = orc_logit_normal_transform(nm1, mu, w->sigma, w->zs, w->ys, ladj_on); /* :522 */
    for (int64_t i = 0; i < nm1; ++i) {                                                /* :523 */
        if (w->ys[i] < eps) w->ys[i] = eps;
        if (w->ys[i] > 1.0 - eps) w->ys[i] = 1.0 - eps;
    }
    double hsb_ladj = orc_ptt_transform(t, w->ys, w->xs, ladj_on);                     /* :525 */
    for (int64_t j = 0; j < n; ++j) { /* :526 clamp!(xs::Vector{Float32}, 1e-10, 1-1e-10) */
        double v = (double)w->xs[j];
        if (v < eps) v = eps;
        if (v > 1.0 - eps) v = 1.0 - eps;
        w->xs[j] = (float)v;
    }
    double lp;
    if (ks)
        lp = orc_factored_log_likelihood(M->m, n, M->frag_probs, M->log_frag_probs, M->colptr,
                                         M->rowval, M->nzval, M->t_colptr, M->t_rowval, M->t_nzval,
                                         ks, w->xs, w->x_grad, lik_gradonly);
    else
        lp = orc_log_likelihood(M->m, n, M->frag_probs, M->log_frag_probs, M->colptr, M->rowval,
                                M->nzval, M->t_colptr, M->t_rowval, M->t_nzval, w->xs, w->x_grad,
                                lik_gradonly);                                         /* :528 */
    if (use_efflen_jacobian)
        lp += orc_effective_length_jacobian_adjustment(n, efflens, w->xs, w->xls, w->x_grad); /* :531 */
    if (genes && !ks)                                                                  /* :535-538 */
        lp += orc_gene_noninformative_prior(n, efflens, w->xls, w->xl_grad, w->xs, w->x_grad, genes);
    double elbo = lp + skew_ladj + ln_ladj + hsb_ladj;                                 /* :540 */

    orc_ptt_transform_gradients(t, w->ys, w->y_grad, w->x_grad);                       /* :542 */
    orc_logit_normal_transform_gradients(nm1, w->zs, w->ys, mu, w->sigma, w->y_grad, w->z_grad,
                                         mu_grad, w->sigma_grad);                      /* :543 */
    orc_sinh_asinh_transform_gradients(nm1, zs0, alpha, w->z_grad, alpha_grad);        /* :544 */
    for (int64_t i = 0; i < nm1; ++i) {                                                /* :547-549 */
        float p = w->sigma[i] * w->sigma_grad[i];
        omega_grad[i] = omega_grad[i] + p;
    }
    return elbo;
}

struct orc_fit_state {
    orc_model M;
    orc_ptt *t;
    draw_ws w;
    const int64_t *ks;
    const float *efflens;
    orc_fit_opts o;
    int64_t n;
    int step; /* steps completed */
    float *buf, *mu, *omega, *alpha;
    double last_elbo;
};

orc_fit_state *orc_fit_begin(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                             const float *nzval, const int64_t *ks, const float *efflens,
                             const int32_t *node_parent_idxs, const int32_t *node_js, const orc_fit_opts *o) {
    orc_fit_state *s = (orc_fit_state *)calloc(1, sizeof(orc_fit_state));
    int64_t nm1 = n - 1;
    model_init(&s->M, m, n, colptr, rowval, nzval);
    s->t = orc_ptt_new(node_parent_idxs, node_js, 2 * n - 1);
    ws_init(&s->w, n);
    s->ks = ks; s->efflens = efflens; s->o = *o; s->n = n; s->step = 0;
    s->buf = (float *)calloc((size_t)nm1 * 13, sizeof(float));
    s->mu = s->buf + 10 * nm1; s->omega = s->buf + 11 * nm1; s->alpha = s->buf + 12 * nm1;
    lsn_init(s->t, n, s->mu, s->omega, s->alpha);
    return s;
}

/* one iteration of the loop at likelihood-approximation.jl:496-575; returns 0 or the failing step */
int orc_fit_step(orc_fit_state *s) {
    int64_t nm1 = s->n - 1;
    float *buf = s->buf;
    float *m_mu = buf, *m_omega = buf + nm1, *m_alpha = buf + 2 * nm1;
    float *v_mu = buf + 3 * nm1, *v_omega = buf + 4 * nm1, *v_alpha = buf + 5 * nm1;
    float *mu_grad = buf + 6 * nm1, *omega_grad = buf + 7 * nm1, *alpha_grad = buf + 8 * nm1;
    float *zs0 = buf + 9 * nm1;
    const double ss_mu = 2e-1, ss_omega = 2e-1, ss_alpha = 2e-2; /* :421-423 */
    const orc_fit_opts *o = &s->o;
    int K = o->num_mc_samples;
    int step = s->step + 1;                                                  /* :496 */
    double lr = orc_adam_learning_rate(step - 1);                            /* :497 */
    double elbo = 0.0;
    memset(mu_grad, 0, sizeof(float) * (size_t)nm1 * 3);                     /* :501-503 */
    for (int64_t i = 0; i < nm1; ++i) s->w.sigma[i] = expf(s->omega[i]);     /* :505-507 */
    for (int d = 0; d < K; ++d) {                                            /* :511 */
        if (o->noise)
            memcpy(zs0, o->noise + ((size_t)(step - 1) * K + d) * (size_t)nm1, sizeof(float) * (size_t)nm1);
        else
            orc_noise_fill(o->seed, step - 1, d, nm1, zs0);                  /* :517-519 */
        double e = lsn_draw_body(&s->M, s->t, s->ks, s->efflens, o->gradonly, o->use_efflen_jacobian, o->genes, s->mu,
                                 s->alpha, zs0, &s->w, mu_grad, omega_grad, alpha_grad);
        if (o->elbo_fix) elbo += e; else elbo = e;                           /* :540 (quirk) */
    }
    int all_finite = 1;
    for (int64_t i = 0; i < nm1; ++i) {                                      /* :552-558 */
        mu_grad[i] /= (float)K;
        omega_grad[i] /= (float)K;
        alpha_grad[i] /= (float)K;
        all_finite &= isfinite(mu_grad[i]) && isfinite(omega_grad[i]) && isfinite(alpha_grad[i]);
    }
    if (!all_finite) return step;                                            /* :559 @assert */
    elbo /= (double)K;                                                       /* :561 */
    s->last_elbo = elbo;
    orc_adam_update_mv(nm1, m_mu, v_mu, mu_grad, step);                      /* :566-568 */
    orc_adam_update_mv(nm1, m_omega, v_omega, omega_grad, step);
    orc_adam_update_mv(nm1, m_alpha, v_alpha, alpha_grad, step);
    orc_adam_update_params(nm1, s->mu, m_mu, v_mu, lr, step, ss_mu);         /* :570-572 */
    orc_adam_update_params(nm1, s->omega, m_omega, v_omega, lr, step, ss_omega);
    orc_adam_update_params(nm1, s->alpha, m_alpha, v_alpha, lr, step, ss_alpha);
    s->step = step;
    return 0;
}

void orc_fit_params(const orc_fit_state *s, float *mu, float *omega, float *alpha) {
    size_t b = sizeof(float) * (size_t)(s->n - 1);
    memcpy(mu, s->mu, b); memcpy(omega, s->omega, b); memcpy(alpha, s->alpha, b);
}

void orc_fit_end(orc_fit_state *s) {
    if (!s) return;
    free(s->buf);
    ws_free(&s->w);
    orc_ptt_free(s->t);
    model_free(&s->M);
    free(s);
}

int orc_fit_lsn_ptt(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                    const float *nzval, const int64_t *ks, const float *efflens,
                    const int32_t *node_parent_idxs, const int32_t *node_js, const orc_fit_opts *o,
                    float *mu, float *omega, float *alpha, double *elbo_traj) {
    orc_fit_state *s = orc_fit_begin(m, n, colptr, rowval, nzval, ks, efflens, node_parent_idxs, node_js, o);
    int status = 0;
    for (int step = 1; step <= o->num_steps && !status; ++step) {
        status = orc_fit_step(s);
        if (!status && elbo_traj) elbo_traj[step - 1] = s->last_elbo;
    }
    orc_fit_params(s, mu, omega, alpha);
    orc_fit_end(s);
    return status;
}

double orc_lsn_draw(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                    const float *nzval, const int64_t *ks, const float *efflens,
                    const int32_t *node_parent_idxs, const int32_t *node_js, int gradonly,
                    int use_efflen_jacobian, const float *mu, const float *omega, const float *alpha,
                    const float *zs0, float *xs, double *ys, double *x_grad, float *y_grad,
                    float *mu_grad, float *omega_grad, float *alpha_grad) {
    return orc_lsn_draw_genes(m, n, colptr, rowval, nzval, ks, efflens, node_parent_idxs, node_js, gradonly,
                              use_efflen_jacobian, NULL, mu, omega, alpha, zs0, xs, ys, x_grad, y_grad, mu_grad,
                              omega_grad, alpha_grad);
}

double orc_lsn_draw_genes(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                          const float *nzval, const int64_t *ks, const float *efflens,
                          const int32_t *node_parent_idxs, const int32_t *node_js, int gradonly,
                          int use_efflen_jacobian, const orc_genes *genes, const float *mu, const float *omega,
                          const float *alpha, const float *zs0, float *xs, double *ys, double *x_grad,
                          float *y_grad, float *mu_grad, float *omega_grad, float *alpha_grad) {
    int64_t nm1 = n - 1;
    orc_model M;
    model_init(&M, m, n, colptr, rowval, nzval);
    orc_ptt *t = orc_ptt_new(node_parent_idxs, node_js, 2 * n - 1);
    draw_ws w;
    ws_init(&w, n);
    for (int64_t i = 0; i < nm1; ++i) w.sigma[i] = expf(omega[i]);
    memset(mu_grad, 0, sizeof(float) * (size_t)nm1);
    memset(omega_grad, 0, sizeof(float) * (size_t)nm1);
    memset(alpha_grad, 0, sizeof(float) * (size_t)nm1);
    double e = lsn_draw_body(&M, t, ks, efflens, gradonly, use_efflen_jacobian, genes, mu, alpha, zs0, &w,
                             mu_grad, omega_grad, alpha_grad);
    memcpy(xs, w.xs, sizeof(float) * (size_t)n);
    memcpy(ys, w.ys, sizeof(double) * (size_t)nm1);
    memcpy(x_grad, w.x_grad, sizeof(double) * (size_t)n);
    memcpy(y_grad, w.y_grad, sizeof(float) * (size_t)nm1);
    ws_free(&w);
    orc_ptt_free(t);
    model_free(&M);
    return e;
}

/* likelihood-approximation.jl:149-242 */
int orc_fit_optimize_ptt(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                         const float *nzval, const float *efflens, int num_steps, float *xs) {
    int64_t nm1 = n - 1, N = 2 * n - 1;
    orc_model M;
    model_init(&M, m, n, colptr, rowval, nzval);
    int32_t *pi = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int32_t *js = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    orc_list_nodes(n, pi, js);                                               /* :160 :sequential */
    orc_ptt *t = orc_ptt_new(pi, js, N);
    float *m_z = (float *)calloc((size_t)nm1, 4), *v_z = (float *)calloc((size_t)nm1, 4);
    float *zs = (float *)calloc((size_t)nm1, 4), *xls = (float *)calloc((size_t)n, 4);
    double *ys = (double *)calloc((size_t)nm1, 8), *z_grad = (double *)calloc((size_t)nm1, 8);
    double *y_grad = (double *)calloc((size_t)nm1, 8), *x_grad = (double *)calloc((size_t)n, 8);
    const double ss_max_z_step = 1e-1, eps = 1e-10;
    for (int64_t i = 0; i < n; ++i) xs[i] = 1.0f / (float)n;                 /* :184-188 */
    orc_ptt_inverse_transform(t, xs, ys);
    for (int64_t i = 0; i < nm1; ++i) zs[i] = (float)log(ys[i] / (1.0 - ys[i]));
    for (int step = 1; step <= num_steps; ++step) {
        double lr = orc_adam_learning_rate(step - 1);
        for (int64_t i = 0; i < nm1; ++i) ys[i] = (double)(1.0f / (1.0f + expf(-zs[i]))); /* :196 */
        memset(x_grad, 0, sizeof(double) * (size_t)n);
        memset(y_grad, 0, sizeof(double) * (size_t)nm1);
        orc_ptt_transform(t, ys, xs, 0);                                     /* :202 */
        for (int64_t j = 0; j < n; ++j) {                                    /* :203 */
            double v = (double)xs[j];
            if (v < eps) v = eps;
            if (v > 1.0 - eps) v = 1.0 - eps;
            xs[j] = (float)v;
        }
        orc_log_likelihood(m, n, M.frag_probs, M.log_frag_probs, colptr, rowval, nzval, M.t_colptr,
                           M.t_rowval, M.t_nzval, xs, x_grad, 1);            /* :205 */
        orc_effective_length_jacobian_adjustment(n, efflens, xs, xls, x_grad); /* :207 */
        orc_ptt_transform_gradients_no_ladj(t, ys, y_grad, x_grad);          /* :209 */
        for (int64_t i = 0; i < nm1; ++i) z_grad[i] = (ys[i] * (1.0 - ys[i])) * y_grad[i]; /* :211-213 */
        /* adam_update_mv! with grad::Vector{Float64}: grad^2 in Float64 */
        if (step == 1) {
            for (int64_t i = 0; i < nm1; ++i) { m_z[i] = (float)z_grad[i]; v_z[i] = (float)(z_grad[i] * z_grad[i]); }
        } else {
            for (int64_t i = 0; i < nm1; ++i) {
                double a = ORC_ADAM_RM * (double)m_z[i], b = (1.0 - ORC_ADAM_RM) * z_grad[i];
                m_z[i] = (float)(a + b);
                double c = ORC_ADAM_RV * (double)v_z[i], d = (1.0 - ORC_ADAM_RV) * (z_grad[i] * z_grad[i]);
                v_z[i] = (float)(c + d);
            }
        }
        orc_adam_update_params(nm1, zs, m_z, v_z, lr, step, ss_max_z_step);  /* :232 */
    }
    for (int64_t i = 0; i < nm1; ++i) ys[i] = (double)(1.0f / (1.0f + expf(-zs[i])));     /* :237-241 */
    orc_ptt_transform(t, ys, xs, 0);
    for (int64_t j = 0; j < n; ++j) {
        double v = (double)xs[j];
        if (v < eps) v = eps;
        if (v > 1.0 - eps) v = 1.0 - eps;
        xs[j] = (float)v;
    }
    free(m_z); free(v_z); free(zs); free(xls); free(ys); free(z_grad); free(y_grad); free(x_grad);
    free(pi); free(js);
    orc_ptt_free(t);
    model_free(&M);
    return 0;
}

/* =============================== hsb_ops.cpp =============================== */

/* hsb_ops.cpp:87-109: index tensors are [B, 2n-1] (a tree per batch row), 0-based, leaf < 0 = internal */
void orc_hsb(int64_t B, int64_t n, const float *y_logit, const int32_t *left, const int32_t *right,
             const int32_t *leaf, float *x) {
    int64_t N = 2 * n - 1;
    double *u = (double *)malloc(sizeof(double) * (size_t)N);
    for (int64_t i = 0; i < B; ++i) {
        const int32_t *L = left + i * N, *R = right + i * N, *F = leaf + i * N;
        u[0] = 1.0;
        int64_t k = 0;
        for (int64_t j = 0; j < N; ++j) {
            if (F[j] >= 0) {
                x[i * n + F[j]] = (float)u[j];
            } else {
                /* `(double) exp(-y_logit_i[k])`: unqualified exp() on a float resolves to ::exp(double)
                 * (verified bit-for-bit against the reference build in oracle/_ref) */
                double y = 1.0 / (1.0 + exp((double)(-y_logit[i * (n - 1) + k])));
                u[L[j]] = y * u[j];
                u[R[j]] = (1.0 - y) * u[j];
                ++k;
            }
        }
    }
    free(u);
}

/* hsb_ops.cpp:206-239; ladj is a float accumulator, log() of a double */
void orc_inv_hsb(int64_t B, int64_t n, const float *x, const int32_t *left, const int32_t *right,
                 const int32_t *leaf, double *y, float *ladj) {
    int64_t N = 2 * n - 1;
    double *u = (double *)malloc(sizeof(double) * (size_t)N);
    for (int64_t i = 0; i < B; ++i) {
        const int32_t *L = left + i * N, *R = right + i * N, *F = leaf + i * N;
        ladj[i] = 0.0f;
        int64_t k = n - 2;
        for (int64_t j = N - 1; j >= 0; --j) {
            if (F[j] >= 0) {
                u[j] = (double)x[i * n + F[j]];
            } else {
                double ul = u[L[j]], ur = u[R[j]];
                u[j] = ul + ur;
                y[i * (n - 1) + k] = ul / u[j];
                ladj[i] = (float)((double)ladj[i] - log(u[j]));
                --k;
            }
        }
    }
    free(u);
}

/* hsb_ops.cpp:338-392 */
void orc_inv_hsb_grad(int64_t B, int64_t n, const double *y_grad, const float *ladj_grad,
                      const double *yv, const int32_t *left, const int32_t *right,
                      const int32_t *leaf, float *backprops) {
    int64_t N = 2 * n - 1;
    double *u = (double *)malloc(sizeof(double) * (size_t)N);
    double *v = (double *)malloc(sizeof(double) * (size_t)N);
    for (int64_t i = 0; i < B; ++i) {
        const int32_t *L = left + i * N, *R = right + i * N, *F = leaf + i * N;
        u[0] = 1.0;
        v[0] = 0.0;
        int64_t k = 0;
        for (int64_t j = 0; j < N; ++j) {
            if (F[j] >= 0) {
                backprops[i * n + F[j]] = (float)v[j];
            } else {
                double y = yv[i * (n - 1) + k];
                double u_j = u[j], u_left = u_j * y, u_right = u_j * (1.0 - y);
                double dladj_du = -1.0 / u_j, u_j2 = u_j * u_j;
                double lg = (double)ladj_grad[i];
                v[L[j]] = (dladj_du * lg + v[j]) + (u_right / u_j2) * y_grad[i * (n - 1) + k];
                v[R[j]] = (dladj_du * lg + v[j]) - (u_left / u_j2) * y_grad[i * (n - 1) + k];
                u[L[j]] = u_left;
                u[R[j]] = u_right;
                ++k;
            }
        }
    }
    free(u);
    free(v);
}
