/*
 * polee_oracle.h -- CPU restatement of dcjones/polee's prep-sample likelihood-approximation path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (libpolee_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * Every function restates one reference function loop-for-loop and dtype-for-dtype
 * (Float32 storage / Float64 accumulators exactly where Julia has them); citations are
 * file:line relative to the reference checkout.  All index arrays are passed exactly as Julia
 * stores them (1-based UInt32 / Int32).
 *
 * Parity status (see DESIGN.md "Oracle pinning"):
 *   - pinned against the reference's one golden pair (test/dataset/mBr_M_6w_1.likelihood-matrix.h5 ->
 *     mBr_M_6w_1.prep.h5) statistically, and against the reference's own hsb_ops.cpp compiled unmodified
 *     over a stub TensorFlow API (oracle/_ref/) for the three HSB ops (bit-level);
 *   - the Julia arithmetic itself cannot run here (no Julia in the image): "parity unpinned" at
 *     the bit level for the Julia functions, and for the RNG stream.
 */
#ifndef POLEE_ORACLE_H
#define POLEE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants.jl:45-65 ---- */
#define ORC_LIKAP_Y_EPS 1e-10
#define ORC_ADAM_INITIAL_LEARNING_RATE 1.0
#define ORC_ADAM_LEARNING_RATE_DECAY 2e-2
#define ORC_ADAM_MIN_LEARNING_RATE 1e-3
#define ORC_ADAM_EPS 1e-8
#define ORC_ADAM_RV 0.9
#define ORC_ADAM_RM 0.7
#define ORC_LIKAP_NUM_STEPS 500
#define ORC_LIKAP_NUM_MC_SAMPLES 6

/* ---- sparse.jl ---- */
/* pAt_mul_B!(y::Vector{Float64}, A::SparseMatrixCSC{Float32,UInt32}, x::Vector{Float32})  sparse.jl:6-21 */
void orc_pAt_mul_B_f32(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, const float *x);
/* same with x::Vector{Float64} (factored path, likelihood.jl:82) */
void orc_pAt_mul_B_f64(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, const double *x);
/* pAt_mulinv_B!(y, A, x::Vector{Float64})  sparse.jl:25-40 */
void orc_pAt_mulinv_B(double *y, int64_t ncols, const uint32_t *colptr, const uint32_t *rowval,
                      const float *nzval, const double *x);
/* Xt = SparseMatrixCSC(transpose(X))  likelihood-approximation.jl:407; outputs 1-based, sorted */
void orc_transpose_csc(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                       const float *nzval, uint32_t *t_colptr /*m+1*/, uint32_t *t_rowval,
                       float *t_nzval);

/* ---- likelihood.jl ---- */
/* log_likelihood  likelihood.jl:36-56; frag_probs/log_frag_probs are Float64[m] work vectors */
double orc_log_likelihood(int64_t m, int64_t n, double *frag_probs, double *log_frag_probs,
                          const uint32_t *colptr, const uint32_t *rowval, const float *nzval,
                          const uint32_t *t_colptr, const uint32_t *t_rowval, const float *t_nzval,
                          const float *xs, double *x_grad, int gradonly);
/* factored_log_likelihood  likelihood.jl:59-85 */
double orc_factored_log_likelihood(int64_t m, int64_t n, double *frag_probs, double *log_frag_probs,
                                   const uint32_t *colptr, const uint32_t *rowval, const float *nzval,
                                   const uint32_t *t_colptr, const uint32_t *t_rowval,
                                   const float *t_nzval, const int64_t *ks, const float *xs,
                                   double *x_grad, int gradonly);
/* effective_length_jacobian_adjustment!  likelihood.jl:93-110 */
double orc_effective_length_jacobian_adjustment(int64_t n, const float *efflens, const float *xs,
                                                float *xls, double *x_grad);

/* ---- ptt.jl ---- */
typedef struct {
    int64_t num_nodes;
    int32_t *index;   /* 4 x num_nodes, column-major, 1-based (ptt.jl:6-12) */
    double *us;       /* num_nodes */
    float *gradients; /* 2 x num_nodes, column-major */
} orc_ptt;

/* PolyaTreeTransform(parent_idxs, output_idxs)  ptt.jl:89-116 */
orc_ptt *orc_ptt_new(const int32_t *parent_idxs, const int32_t *js, int64_t num_nodes);
void orc_ptt_free(orc_ptt *t);
const int32_t *orc_ptt_index(const orc_ptt *t);
/* transform!  ptt.jl:125-160 (xs Float32) */
double orc_ptt_transform(orc_ptt *t, const double *ys, float *xs, int compute_ladj);
/* transform_gradients!  ptt.jl:167-209 (y_grad Float32, x_grad Float64) */
void orc_ptt_transform_gradients(orc_ptt *t, const double *ys, float *y_grad, const double *x_grad);
/* transform_gradients_no_ladj!  ptt.jl:217-251 (y_grad Float64 as in OptimizePTTApprox) */
void orc_ptt_transform_gradients_no_ladj(orc_ptt *t, const double *ys, double *y_grad,
                                         const double *x_grad);
/* inverse_transform!  ptt.jl:257-285 (xs Float32, ys Float64) */
double orc_ptt_inverse_transform(orc_ptt *t, const float *xs, double *ys);
/* make_inverse_ptt_params  ptt.jl:293-309 */
void orc_make_inverse_ptt_params(const int32_t *node_parent_idxs, const int32_t *node_js,
                                 int64_t num_nodes, int32_t *left_index, int32_t *right_index,
                                 int32_t *leaf_index);
/* list_nodes / order_nodes  hclust.jl:477-489, 361-389 -> flattened (parent_idxs, js) */
void orc_list_nodes(int64_t n, int32_t *parent_idxs, int32_t *js);

/* ---- logitnormal.jl / sinh_arcsinh.jl ---- */
double orc_sinh_asinh_transform(int64_t nm1, const float *alpha, const float *zs0, float *zs,
                                int compute_ladj);
double orc_logit_normal_transform(int64_t nm1, const float *mu, const float *sigma, const float *zs,
                                  double *ys, int compute_ladj);
void orc_logit_normal_transform_gradients(int64_t nm1, const float *zs, const double *ys,
                                          const float *mu, const float *sigma, const float *y_grad,
                                          float *z_grad, float *mu_grad, float *sigma_grad);
void orc_sinh_asinh_transform_gradients(int64_t nm1, const float *zs0, const float *alpha,
                                        const float *z_grad, float *alpha_grad);

/* ---- likelihood-approximation.jl:107-146 ---- */
double orc_adam_learning_rate(int64_t step_num);
void orc_adam_update_mv(int64_t len, float *ms, float *vs, const float *grad, int64_t step_num);
void orc_adam_update_params(int64_t len, float *params, const float *ms, const float *vs,
                            double learning_rate, int64_t step_num, double max_step_size);

/* ---- noise: Philox4x32-10 + Box-Muller (this project's spec "polee-philox-v1"; the
 *      reference uses Julia's global MersenneTwister randn(Float32), which cannot be reproduced) ---- */
void orc_noise_fill(uint64_t seed, int64_t step, int64_t draw, int64_t nm1, float *zs0);

/* gene -> transcripts map of gene_noninformative_prior! (likelihood-approximation.jl:476-493), flattened */
typedef struct {
    int64_t num_genes;
    const int64_t *gene_ptr;     /* num_genes + 1 offsets into transcripts */
    const int32_t *transcripts;  /* 1-based transcript ids */
} orc_genes;

/* likelihood.jl:114-159 (xls as left by orc_effective_length_jacobian_adjustment) */
double orc_gene_noninformative_prior(int64_t n, const float *efflens, const float *xls, double *xl_grad,
                                     const float *xs, double *x_grad, const orc_genes *genes);

/* ---- the fits ---- */
typedef struct {
    int num_steps;            /* LIKAP_NUM_STEPS */
    int num_mc_samples;       /* LIKAP_NUM_MC_SAMPLES */
    int gradonly;             /* Val(gradonly), default 1 */
    int use_efflen_jacobian;  /* default 1 */
    uint64_t seed;
    const float *noise;       /* nullable: [num_steps][num_mc_samples][n-1] injected zs0 */
    int elbo_fix;             /* 0: reference quirk (elbo assigned per draw); 1: mean over draws */
    const orc_genes *genes;   /* nullable: gene_noninformative=true (likelihood-approximation.jl:535-538) */
} orc_fit_opts;

/* approximate_likelihood(::LogitSkewNormalPTTApprox, sample)  likelihood-approximation.jl:395-624
 * (ks == NULL) and the factored variant :248-392 (ks != NULL).  Outputs mu/omega/alpha (n-1) and
 * optionally the per-step elbo (num_steps).  Returns 0, or the 1-based step with a non-finite gradient. */
int orc_fit_lsn_ptt(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                    const float *nzval, const int64_t *ks, const float *efflens,
                    const int32_t *node_parent_idxs, const int32_t *node_js, const orc_fit_opts *opts,
                    float *mu, float *omega, float *alpha, double *elbo_traj);

/* the same fit, one ADAM step at a time (bench.py's CPU baseline times orc_fit_step) */
typedef struct orc_fit_state orc_fit_state;
orc_fit_state *orc_fit_begin(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                             const float *nzval, const int64_t *ks, const float *efflens,
                             const int32_t *node_parent_idxs, const int32_t *node_js, const orc_fit_opts *opts);
int orc_fit_step(orc_fit_state *s);
void orc_fit_params(const orc_fit_state *s, float *mu, float *omega, float *alpha);
void orc_fit_end(orc_fit_state *s);

/* one MC draw of the loop body (:512-549) at fixed parameters and injected noise: outputs the
 * per-draw gradient contributions and intermediates for piecewise parity tests. */
double orc_lsn_draw(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                    const float *nzval, const int64_t *ks, const float *efflens,
                    const int32_t *node_parent_idxs, const int32_t *node_js, int gradonly,
                    int use_efflen_jacobian, const float *mu, const float *omega, const float *alpha,
                    const float *zs0, float *xs /*n*/, double *ys /*n-1*/, double *x_grad /*n*/,
                    float *y_grad /*n-1*/, float *mu_grad, float *omega_grad, float *alpha_grad);

/* the same with the gene prior switched on (genes nullable) */
double orc_lsn_draw_genes(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                          const float *nzval, const int64_t *ks, const float *efflens,
                          const int32_t *node_parent_idxs, const int32_t *node_js, int gradonly,
                          int use_efflen_jacobian, const orc_genes *genes, const float *mu, const float *omega,
                          const float *alpha, const float *zs0, float *xs, double *ys, double *x_grad,
                          float *y_grad, float *mu_grad, float *omega_grad, float *alpha_grad);

/* approximate_likelihood(::OptimizePTTApprox, sample)  likelihood-approximation.jl:149-242 */
int orc_fit_optimize_ptt(int64_t m, int64_t n, const uint32_t *colptr, const uint32_t *rowval,
                         const float *nzval, const float *efflens, int num_steps, float *xs_out);

/* ---- hsb_ops.cpp (restated on raw pointers) ---- */
void orc_hsb(int64_t B, int64_t n, const float *y_logit, const int32_t *left, const int32_t *right,
             const int32_t *leaf, float *x);                                  /* hsb_ops.cpp:87-109 */
void orc_inv_hsb(int64_t B, int64_t n, const float *x, const int32_t *left, const int32_t *right,
                 const int32_t *leaf, double *y, float *ladj);                /* hsb_ops.cpp:206-239 */
void orc_inv_hsb_grad(int64_t B, int64_t n, const double *y_grad, const float *ladj_grad,
                      const double *y, const int32_t *left, const int32_t *right,
                      const int32_t *leaf, float *backprops);                 /* hsb_ops.cpp:338-392 */

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
