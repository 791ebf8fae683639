/* sqair_b200 -- C ABI of the B200-native SQAIR Discover/Propagate hot path.
 *
 * The reference (akosiorek/sqair @ 474f5d0) is pure Python building a TF1/Sonnet graph; it has no
 * FFI.  The operator this library replaces is `SequentialAIR.__call__` (sqair/seq.py:69-84: the
 * `tf.while_loop` over frames whose body is seq.py:181-269) together with the objective of
 * `Model._build` / `Model.make_target` (sqair/model.py:79-168, sqair/targets.py:38-75).  Each entry
 * point below names the reference code it stands in for.  The ctypes binding a maintainer of the
 * reference would add is shown in INTEGRATION.md.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller unless the name says `host`;
 * the library never allocates device memory (except one small per-configuration descriptor table that it caches for the
 * process lifetime); all work is enqueued on `stream` (a cudaStream_t passed
 * as void*; NULL = legacy default stream) and is asynchronous.  Functions return 0 on success or a
 * negative code; `sqair_last_error()` returns a thread-local message for the last failure.
 * All tensors are fp32, C-contiguous.
 */
#ifndef SQAIR_B200_H
#define SQAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SQAIR_OK = 0, SQAIR_EINVAL = -1, SQAIR_EUNSUPPORTED = -2, SQAIR_ECUDA = -3 };
enum { SQAIR_PRIOR_RNN = 0, SQAIR_PRIOR_RW = 1, SQAIR_PRIOR_GUIDED = 2 };   /* propagate.py:35-45 */
enum { SQAIR_DISC_PRIOR_CAT = 0, SQAIR_DISC_PRIOR_GEOM = 1 };               /* sqair_modules.py:205-224 */

/* Model / problem description: the flags of common_model_flags.py:32-56 and
 * configs/mlp_mnist_model.py:42-52 plus the tensor sizes of the call. */
typedef struct sqair_cfg {
    int32_t T;                 /* frames in the sequence                                   */
    int32_t B;                 /* sequences handled by THIS call (the local shard)          */
    int32_t K;                 /* k_particles; rows = B*K, row = b*K + k (index.py:106-129) */
    int32_t n;                 /* n_steps_per_image (object slots)                          */
    int32_t H, W;              /* canvas size (C = 1)                                       */
    int32_t G;                 /* glimpse_size                                              */
    int32_t n_what;            /* n_what                                                    */
    int32_t n_hidden;          /* 32 * n_units                                              */
    int32_t prior_type;        /* SQAIR_PRIOR_*                                             */
    int32_t disc_prior_type;   /* SQAIR_DISC_PRIOR_*                                        */
    int32_t rec_where_prior;   /* bool                                                      */
    int32_t masked_glimpse;    /* bool                                                      */
    float step_success_prob;   /* geom prior only                                           */
    float prop_prior_step_bias;
    float output_std, bg_std;  /* std of p(x|z) inside / outside glimpses (modules.py:406-426) */
    float where_update_scale;  /* core.py:258-260 (non-trainable)                           */
    float min_std;             /* modules.py:72 (non-trainable)                             */
    float where_mean[4], where_std[4];   /* disc where prior when rec_where_prior == 0      */
} sqair_cfg;

typedef struct sqair_sizes {
    int64_t param_count;       /* trainable scalars, canonical (TF variable order) layout   */
    int64_t packed_floats;     /* floats of the kernel-side parameter buffer                */
    int64_t eps_where_floats, eps_what_floats, u_pres_floats;   /* noise tensors            */
    int32_t rows;              /* B*K                                                       */
    int32_t rows_per_cta;      /* rows each cluster of thread blocks carries through the sequence */
    int32_t cluster_size;      /* thread blocks per cluster (column split of every dense layer) */
    int32_t n_ctas;
    int32_t smem_bytes;        /* dynamic shared memory per thread block                    */
    int32_t n_layers;          /* dense-layer calls per frame                               */
} sqair_sizes;

/* One trainable variable of the reference (names as printed by notebooks/play.ipynb:239-362). */
typedef struct sqair_param_desc {
    char name[160];
    int32_t ndim;
    int32_t shape[3];
    int64_t offset;            /* into the canonical flat buffer (tight, variable order)    */
    int64_t packed_offset;     /* into the kernel-side buffer (16-byte aligned)             */
} sqair_param_desc;

/* The 38 per-frame outputs of SequentialAIR (seq.py:121-177), each [T, rows, ...].  NULL = skip. */
typedef struct sqair_outputs {
    float *what, *what_loc, *what_scale;          /* [T,rows,n,n_what]                       */
    float *where, *where_loc, *where_scale;       /* [T,rows,n,4]                            */
    float *presence_prob, *presence, *presence_logit;   /* [T,rows,n]                        */
    float *obj_id;                                /* [T,rows,n]                              */
    float *step_log_prob;                         /* [T,rows]                                */
    float *canvas;                                /* [T,rows,H,W]                            */
    float *glimpse;                               /* [T,rows,n,G,G]                          */
    float *disc_what_log_prob, *disc_where_log_prob, *disc_what_prior_log_prob, *disc_where_prior_log_prob; /* [T,rows,n] */
    float *disc_log_prob, *disc_prior_log_prob;   /* [T,rows]                                */
    float *disc_prob;                             /* [T,rows,n+1]                            */
    float *prop_what_log_prob, *prop_where_log_prob, *prop_what_prior_log_prob, *prop_where_prior_log_prob; /* [T,rows,n] */
    float *prop_log_prob, *prop_prior_log_prob;   /* [T,rows]                                */
    float *prop_prob;                             /* [T,rows,n]                              */
    float *discrete_log_prob;                     /* [T,rows]                                */
    float *num_prop_steps_per_sample, *num_disc_steps_per_sample, *num_steps_per_sample;   /* [T,rows] */
    float *prop_pres, *disc_pres;                 /* [T,rows,n]                              */
    float *data_ll_per_sample, *kl_per_sample, *log_q_z_given_x_per_sample, *log_p_z_per_sample,
          *log_weights_per_timestep;              /* [T,rows]                                */
} sqair_outputs;

/* Scalars produced by sqair_objective (model.py:88-103,150-158; targets.py:38-75; ops.py:52-59). */
enum { SQAIR_OBJ_ELBO_VAE = 0, SQAIR_OBJ_ELBO_IWAE = 1, SQAIR_OBJ_ESS = 2, SQAIR_OBJ_VIMCO_TARGET = 3,
       SQAIR_OBJ_IWAE_TARGET = 4, SQAIR_OBJ_N = 8 };

const char* sqair_last_error(void);
int sqair_version(void);

/* Sizes implied by a configuration; also validates it (replaces the graph-build shape logic of
 * seq.py:86-179 and the variable creation of configs/mlp_mnist_model.py:74-150). */
int sqair_query_sizes(const sqair_cfg* cfg, sqair_sizes* out);

/* Table of the reference's trainable variables in canonical order.  `descs` may be NULL to query
 * the count; otherwise *n holds the capacity on entry and the count on return. */
int sqair_param_layout(const sqair_cfg* cfg, sqair_param_desc* descs, int32_t* n);

/* canonical flat parameters -> kernel-side (aligned) buffer. */
int sqair_pack_params(const sqair_cfg* cfg, const float* params, float* packed, void* stream);

/* Counter-based draws for every noise site of the path (core.py:143,218,227,
 * 330,356): Philox4x32-10 keyed by `seed`, counter (row_offset + row, frame, slot, block), so a
 * shard of rows reproduces exactly the draws of the unsharded run.
 * eps_where [T,rows,2n,4], eps_what [T,rows,2n,n_what], u_pres [T,rows,2n]. */
int sqair_fill_noise(const sqair_cfg* cfg, uint64_t seed, int32_t row_offset,
                     float* eps_where, float* eps_what, float* u_pres, void* stream);

/* SequentialAIR.__call__ (seq.py:69-84): the whole T-frame Discover/Propagate recursion in one
 * persistent kernel.  obs is [T,B,H,W] (NOT tiled: the K particles of a sequence share its frame,
 * replacing index.tile_input_for_iwae, index.py:106-129). */
int sqair_forward(const sqair_cfg* cfg, const float* packed_params, const float* obs,
                  const float* eps_where, const float* eps_what, const float* u_pres,
                  const sqair_outputs* out, void* stream);

/* Model._build / make_target reductions over particles (model.py:88-103,150-158).
 * log_w_t, disc_lp_t: [T, B*K] per-frame log weights / discrete log-probs.  Outputs (nullable):
 * log_weights [B,K], elbo_iwae_per_example [B], importance_weights [B,K], scalars [SQAIR_OBJ_N]. */
int sqair_objective(const float* log_w_t, const float* disc_lp_t, int32_t T, int32_t B, int32_t K,
                    float* log_weights, float* elbo_iwae_per_example, float* importance_weights,
                    float* scalars, void* stream);

/* First stage of the backward pass of Model.make_target (model.py:150-158; targets.py:46-75): gradients of the
 * VIMCO target `mean_{b,k}(-elbo_iwae_b - stop_gradient(log_w_bk - control_variate_bk) * log_prob_bk) / T` with
 * respect to the summed log weight and the summed discrete log-probability of every row (the same value flows to
 * each of the row's T per-frame terms):
 *   d_log_weights [B,K]        = -softmax_k(log_w_b)_k / (B * T)
 *   d_discrete_log_prob [B,K]  = -(log_w_bk - control_variate_bk) / (B * K * T)
 * Either output may be NULL. */
int sqair_objective_grad(const float* log_w_t, const float* disc_lp_t, int32_t T, int32_t B, int32_t K,
                         float* d_log_weights, float* d_discrete_log_prob, void* stream);

/* ---- training step: forward with stash, backward, gradient in the reference's variable layout -------------------
 * Stand-in for `opt.compute_gradients(target)` of Model.make_target (sqair/model.py:150-168): TF differentiates the
 * whole graph of seq.py:181-276; here the forward kernel additionally writes every activation the adjoint needs into a
 * caller-owned `stash`, and sqair_backward walks the frames in reverse.  Sizes: */
typedef struct sqair_train_sizes {
    int64_t stash_floats;            /* activations written by sqair_forward_train, read by sqair_backward     */
    int64_t workspace_floats;        /* scratch of sqair_backward (per-layer output gradients, state adjoints)   */
    int64_t backward_param_floats;   /* parameter copy in the layout of the backward GEMMs (sqair_pack_backward) */
} sqair_train_sizes;
int sqair_query_train_sizes(const sqair_cfg* cfg, sqair_train_sizes* out);

/* sqair_forward that also fills `stash` (stash_floats floats).  Same outputs, same arithmetic. */
int sqair_forward_train(const sqair_cfg* cfg, const float* packed_params, const float* obs,
                        const float* eps_where, const float* eps_what, const float* u_pres,
                        const sqair_outputs* out, float* stash, void* stream);

/* Generation (seq.py:46,198-203; sqair_modules.py:157-170,294-302): `SequentialAIR(..., sample_from_prior=True,
 * generate_after=g)`.  The inference networks still run on every frame; in addition every propagated slot draws
 * (what, where, presence) from the propagation prior with the second noise set (same shapes as the first; slots
 * 0 .. n-1 are used), the posterior log-probabilities are evaluated at those draws, and in frames t > g (only when
 * g > 0) the draws replace the posterior samples while discovery adds no objects -- the model rolls forward from its
 * prior.  Same outputs as sqair_forward. */
int sqair_forward_generate(const sqair_cfg* cfg, const float* packed_params, const float* obs,
                           const float* eps_where, const float* eps_what, const float* u_pres,
                           const float* eps_where_prior, const float* eps_what_prior, const float* u_pres_prior,
                           int32_t generate_after, const sqair_outputs* out, void* stream);

/* canonical flat parameters -> backward parameter buffer (once per parameter update, like sqair_pack_params): every layer's
 * virtual matrix row-major and transposed (launch-per-operation path, weight-gradient scatter) and W^T as tensor-core
 * fragment panels for the reverse-program kernel; sqair_query_train_sizes gives the size. */
int sqair_pack_backward(const sqair_cfg* cfg, const float* params, float* bw_params, void* stream);

/* Gradient of the training target w.r.t. every variable of sqair_param_layout (canonical flat layout, overwritten).
 * d_log_weights / d_discrete_log_prob: [B*K] from sqair_objective_grad (d_discrete_log_prob may be NULL = zeros: the
 * `-elbo_iwae` target of model.py:156).  `params` is the canonical flat buffer the packed copies were made from; obs and
 * the noise tensors are the ones the forward call consumed.  *n_launches (nullable) receives the number of kernels and
 * memsets enqueued.  The frame recursion runs as ONE persistent cluster kernel interpreting an operation table that the
 * library records and caches per (configuration, buffer addresses); the first call with new buffers uploads the table and
 * synchronises `stream` once.  Otherwise asynchronous on `stream` (the weight-gradient GEMMs fork onto library-owned side
 * streams and join back before the call returns control of `stream`); capturable in a CUDA graph after one eager call with
 * the same buffers.  SQAIR_BWD_LAUNCHES=1 selects the launch-per-operation path instead. */
int sqair_backward(const sqair_cfg* cfg, const float* params, const float* bw_params, const float* obs,
                   const float* eps_where, const float* eps_what, const float* stash,
                   const float* d_log_weights, const float* d_discrete_log_prob,
                   float* workspace, float* d_params, int32_t* n_launches, void* stream);

/* Optimiser step on the flat canonical buffers: `opt.apply_gradients(gvs)` of the reference's training loop
 * (scripts/experiment.py:138-146,153-155) with TensorFlow 1.x update rules.  kind = SQAIR_OPT_*:
 *   RMSPROP  (hyper_a = decay 0.9, hyper_b = momentum 0.9, epsilon 1e-10; slot0 = mean square, initialised to ONE as
 *            tf.train.RMSPropOptimizer does, slot1 = momentum, zero),
 *   ADAM     (hyper_a = beta1, hyper_b = beta2, epsilon 1e-8; `lr` must already carry the bias correction
 *            sqrt(1 - beta2^t) / (1 - beta1^t); slots m, v zero),
 *   MOMENTUM (hyper_a = momentum; slot0 = accumulator), SGD (no slots).
 * The gradient used is grad_scale * grad + l2_weight * param (targets.py:31-35 l2_reg).  In place, asynchronous. */
enum { SQAIR_OPT_RMSPROP = 0, SQAIR_OPT_ADAM = 1, SQAIR_OPT_MOMENTUM = 2, SQAIR_OPT_SGD = 3 };
int sqair_optimizer_update(int32_t kind, float* params, const float* grad, float* slot0, float* slot1, int64_t n, float lr,
                           float hyper_a, float hyper_b, float epsilon, float grad_scale, float l2_weight, void* stream);

/* Data path: renders frames of moving sprites on the device -- `TemplateDataset.create`
 * (data/template.py:58-104: every object pasted at its rounded position, max blend) followed by the uint8 -> float / 255
 * of data/data.py:199.  atlas [S][cell][cell] uint8 templates (top-left aligned), atlas_hw [S][2] their (height, width),
 * pos [T][B][n][2] int32 (y, x) of the top-left corner, sprite [B][n] template index or -1 for "no object",
 * frames [T][B][H][W] float32 (overwritten). */
int sqair_render_sprites(const uint8_t* atlas, const int32_t* atlas_hw, const int32_t* pos, const int32_t* sprite,
                         float* frames, int32_t T, int32_t B, int32_t n, int32_t H, int32_t W, int32_t S, int32_t cell,
                         void* stream);

/* Weight gradient of one dense layer (the GEMM-shaped part of the backward pass): dW [K,N] (+)= X^T dY for
 * the stashed layer inputs X [M,K] and output gradients dY [M,N], M = rows x frames x slots, all row-major fp32.
 * fp32-faithful on the tensor cores (tf32 hi/lo split of both operands, four products, per-k-step fp32 accumulation).
 * accumulate != 0 adds into dW (which must then be initialised), otherwise dW is overwritten. */
int sqair_wgrad(const float* x, const float* dy, float* dw, int32_t M, int32_t K, int32_t N, int32_t accumulate,
                void* stream);

/* Input gradient of one dense layer, the other GEMM of the backward pass: dX [M,K] = dY [M,N] . W^T for the layer's
 * weight matrix W [K,N] in the reference's own (canonical, row-major) layout.  Same fp32-faithful tensor-core
 * arithmetic as sqair_wgrad.  (The fused backward kernel will do this product on transposed fragment panels inside
 * the cluster; this entry point is its unit-parity reference and the building block of an unfused backward.) */
int sqair_dgrad(const float* dy, const float* w, float* dx, int32_t M, int32_t K, int32_t N, void* stream);

/* Per-op entry points (unit parity / roofline of the bandwidth-shaped pieces).
 * sqair_stn_glimpse: SpatialTransformer forward (modules.py:165-172,204-218): img [N,H,W],
 *   where-logits [N,4] -> glimpse [N,G,G].
 * sqair_canvas_ll: AIRDecoder._decode/_add_mean_image + pixel likelihood (modules.py:435-467,
 *   seq.py:272-273): glimpse [N,n,G,G], where [N,n,4], presence [N,n], mean_img [H,W], img [N,H,W]
 *   -> canvas [N,H,W], data_ll [N]. */
int sqair_stn_glimpse(const float* img, const float* where, float* glimpse,
                      int32_t N, int32_t H, int32_t W, int32_t G, void* stream);
/* Backward of sqair_stn_glimpse with respect to the where-logits (the frame is an input and gets no gradient,
 * SURVEY Appendix F): d_glimpse [N,G,G] -> d_where [N,4].  Matches tf.contrib.resampler's warp gradient (bilinear
 * weights differentiated, zero outside the frame) chained through AffineGridWarper, to_coords (sigmoid / tanh,
 * modules.py:220-227) and the straight-through clip of the scale (ops.py:33-42, modules.py:206). */
int sqair_stn_glimpse_grad(const float* img, const float* where, const float* d_glimpse, float* d_where,
                           int32_t N, int32_t H, int32_t W, int32_t G, void* stream);
int sqair_canvas_ll(const float* glimpse, const float* where, const float* presence, const float* mean_img,
                    const float* img, float* canvas, float* data_ll,
                    int32_t N, int32_t n, int32_t H, int32_t W, int32_t G, float output_std, float bg_std,
                    void* stream);

/* Backward of sqair_canvas_ll (AIRDecoder._decode/_add_mean_image + the pixel likelihood, modules.py:435-467,
 * seq.py:272-273) for an upstream gradient d_ll [N] on data_ll: gradients w.r.t. the decoded glimpses
 * d_glimpse [N,n,G,G], the where-logits d_where [N,n,4] (inverse-warp gradient for both the glimpse and the
 * all-ones occupancy map that feeds the mean-image mask) and the trainable mean image d_mean_img [H,W]
 * (ACCUMULATED into the buffer: zero it first).  presence and the frame get no gradient. */
int sqair_canvas_ll_grad(const float* glimpse, const float* where, const float* presence, const float* mean_img,
                         const float* img, const float* d_ll, float* d_glimpse, float* d_where, float* d_mean_img,
                         int32_t N, int32_t n, int32_t H, int32_t W, int32_t G, float output_std, float bg_std,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SQAIR_B200_H */
