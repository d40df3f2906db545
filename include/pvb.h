/*
 * pvb.h -- C ABI of libpvb.so: the hand-written sm_100a kernels behind
 * pyroved_b200 (B200-native SVI hot path of pyroVED).
 *
 * The reference (pure Python, /root/reference/pyroved) has NO FFI/operator
 * interface for this path: its boundary is the Python class
 * API.  The entry points below are therefore the operators a maintainer
 * would bind from the reference's Python via ctypes (INTEGRATION.md shows the
 * stubs); each cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer to contiguous fp32 (unless noted)
 *   - weights use torch's nn.Linear layout  W[out][in]  row-major
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it
 *     and the call never synchronises, allocates or keeps state (re-entrant)
 *   - return value: 0 ok, <0 bad argument (see pvb_last_error_string),
 *     >0 a cudaError_t
 *   - "instance" i in [0, I): one (enumeration index k, sample b) pair,
 *     i = k*B + b ; I = B for iVAE, I = K*B for enumerated jiVAE / ssiVAE
 *     (reference models/jivae.py:182-194, models/ssivae.py:217-227)
 *   - "row" r = i*N + p : pixel p of instance i through the spatial decoder
 */
#ifndef PVB_H_
#define PVB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation codes (reference utils/nn.py:118-124, torch module defaults) */
enum {
  PVB_ACT_NONE = 0,
  PVB_ACT_TANH = 1,
  PVB_ACT_RELU = 2,
  PVB_ACT_LRELU = 3,    /* slope 0.01 */
  PVB_ACT_SOFTPLUS = 4, /* beta 1, threshold 20 */
  PVB_ACT_GELU = 5,     /* erf form */
  PVB_ACT_SIGMOID = 6
};

/* decoder samplers (reference utils/prob.py:25-29) */
enum {
  PVB_SAMPLER_BERNOULLI = 0,
  PVB_SAMPLER_GAUSSIAN = 1,
  PVB_SAMPLER_CONT_BERNOULLI = 2
};

/* invariance flags (reference models/base.py:54-67, 107-118) */
enum { PVB_INV_R = 1, PVB_INV_T = 2, PVB_INV_S = 4 };

int pvb_version(void);
const char* pvb_last_error_string(void);
/* number of kernels this library has launched in this process (statistics) */
long long pvb_launch_count(void);
/* 1 if the library was built with the tcgen05 spatial-decoder kernel */
int pvb_has_tcgen05(void);

/* ---- dense layers: nn.Linear + activation (nets/fc.py:55-61,307-324) ---- */
/* y[M,N] = act(x[M,K] W[N,K]^T + b[N]);  pre (optional, may be NULL) receives
 * the pre-activation (only needed for PVB_ACT_GELU backward).
 * Arithmetic: fp32 throughout.  The small-batch kernel (48..591 tiles of 32 x 32 / 32 x 16, K % 4 == 0,
 * 16-byte aligned rows) and pvb_mlp_wgrad form their products on the warp-level tensor-core path
 * with each fp32 operand split into two TF32 terms (three MMAs per product: error ~2^-22, i.e.
 * fp32-grade); <= 8 outputs over K >= 2048 run on one-pass kernels (forward: a cluster of 8 CTAs
 * along K, partial sums joined in rank order -- deterministic). */
int pvb_linear_fwd(const float* x, const float* W, const float* b, float* y,
                   float* pre, int64_t M, int N, int K, int act, void* stream);
/* Given dy (gradient wrt y), the saved output y (and pre for gelu):
 *   dpre = dy * act'(.)            (written to dpre_ws [M,N], may alias dy)
 *   dx[M,K]  = dpre W              (skipped if dx == NULL; dx_accumulate adds)
 *   dW[N,K] += dpre^T x ; db[N] += colsum(dpre)   (always accumulate) */
int pvb_linear_bwd(const float* x, const float* W, const float* y,
                   const float* pre, const float* dy, float* dpre_ws,
                   float* dx, int dx_accumulate, float* dW, float* db,
                   int64_t M, int N, int K, int act, void* stream);

/* ---- latent sample + sampled KL (models/ivae.py:182-183,217-221) ---- */
/* eps[n] ~ N(0,1): Philox4x32-10 + Box-Muller, element e gets counter
 * (first_index + e, *step_counter), key = seed: identical for any sharding. */
int pvb_randn(float* eps, int64_t n, uint64_t seed, const int32_t* step_counter,
              int64_t first_index, void* stream);
/* sigma = softplus(s_pre); z = mu + sigma*eps;
 * kl[i] = sum_d(-z^2/2 + eps^2/2 + log sigma) = log p(z_i) - log q(z_i). */
int pvb_latent_fwd(const float* mu, const float* s_pre, const float* eps,
                   float* sigma, float* z, float* kl, int64_t I, int Z,
                   void* stream);
/* loss = -sum_i w_i (ll_i + beta kl_i) [+ ...]; gz = dloss/dz through the
 * decoder.  Writes dloss/dmu and dloss/ds_pre.  w may be NULL (all ones). */
int pvb_latent_bwd(const float* gz, const float* eps, const float* sigma,
                   const float* s_pre, const float* z, const float* w,
                   float beta, float* gmu, float* gs_pre, int64_t I, int Z,
                   void* stream);

/* ---- coordinate transform folded into the first decoder layer ----------
 * Replaces split_latent + transform_coordinates + coord_latent pre-activation
 * (models/base.py:97-119, models/ivae.py:187-192, utils/coord.py:47-88,
 * nets/fc.py:226-235; algebra: SURVEY.md Appendix C):
 *   pre0[i,p,:] = Uv[i,0,:]*gx_p + Uv[i,1,:]*gy_p + Uv[i,2,:]
 * z [I,Z] holds (phi | dx dy | s | content[L]) in that fixed order;
 * cond [I,C] (one-hot / class vector concatenated to the content code, may be
 * NULL when C == 0).  Wc [Hd, ndim], bc [Hd], Wz [Hd, L+C] (no bias). */
typedef struct {
  int32_t ndim;        /* 1 or 2 */
  int32_t inv;         /* PVB_INV_* flags (1-D: PVB_INV_T only) */
  int32_t latent_dim;  /* L */
  int32_t cond_dim;    /* C */
  int32_t hidden;      /* Hd */
  float dx_prior, dy_prior, sc_prior;
} pvb_fold_cfg;
int pvb_fold_fwd(const pvb_fold_cfg* cfg, const float* z, const float* cond,
                 const float* Wc, const float* bc, const float* Wz, float* Uv,
                 int64_t I, void* stream);
/* gUv [I,3,Hd] -> gz [I,Z] (overwritten), gcond [I,C] (optional, overwritten),
 * and weight-gradient partials: part[G][Hd*(ndim+1+L+C)] laid out as
 * (gWc[Hd][ndim] | gbc[Hd] | gWz[Hd][L+C]); reduce with pvb_reduce_partials.
 * G = pvb_fold_bwd_num_partials(). */
int pvb_fold_bwd_num_partials(void);
int pvb_fold_bwd(const pvb_fold_cfg* cfg, const float* z, const float* cond,
                 const float* Wc, const float* Wz, const float* gUv, float* gz,
                 float* gcond, float* part, int64_t I, void* stream);

/* ---- spatial decoder, generic fp32 path (any hidden sizes / activation) --
 * h0[r,:] = tanh(pre0) with the grid regenerated from the pixel index
 * (utils/coord.py:14-18,43; nets/fc.py:216-218,236-237 -- always tanh). */
int pvb_sdec_h0_fwd(const float* Uv, float* h0, int64_t I, int H, int W,
                    int ndim, int Hd, void* stream);
/* dh0 [R,Hd] (gradient wrt h0) + saved h0 -> gUv [I,3,Hd] */
int pvb_sdec_h0_bwd(const float* dh0, const float* h0, float* gUv, int64_t I,
                    int H, int W, int ndim, int Hd, void* stream);

/* ---- observation log-likelihood + ELBO reduction -----------------------
 * logit[R] = decoder `out` layer pre-activation, x [B,N] targets, row r of
 * instance i scores against x[i % B] (utils/prob.py:25-29; torch Bernoulli /
 * Normal log_prob; models/ivae.py:200-202).
 *   rowll[r]  = log p(x | loc)              (per pixel)
 *   dlogit[r] = w_i * d(-rowll)/dlogit      (seed of the backward pass)
 *   loc[r]    = reconstruction (sigmoid(logit) if sigmoid_d else logit); may be NULL
 * w may be NULL (all ones). */
int pvb_obs_loglik(const float* logit, const float* x, const float* w,
                   float* rowll, float* dlogit, float* loc, int64_t I,
                   int64_t B, int N, int sampler, int sigmoid_d,
                   float decoder_sig, void* stream);
/* ll[i] = sum_p rowll[i,p] (deterministic warp-shuffle tree);
 * loss_out[0] (+)= -sum_i w_i (ll_i + beta kl_i);  kl, w may be NULL. */
int pvb_elbo_reduce(const float* rowll, const float* kl, const float* w,
                    float beta, float* ll, float* loss_out, int accumulate,
                    int64_t I, int N, void* stream);

/* loss_out[0] += scale * sum_i w_i v_i   (w may be NULL; fixed order) */
int pvb_weighted_sum(const float* v, const float* w, float scale,
                     float* loss_out, int64_t n, void* stream);
/* out[i] = a[i] + beta * b[i] */
int pvb_axpy_out(const float* a, const float* b, float beta, float* out,
                 int64_t n, void* stream);

/* ---- enumerated discrete latent (TraceEnum_ELBO expectation) -----------
 * alpha = softmax(logits [B,K]); w[k*B+b] = alpha[b,k]
 * (nets/fc.py:101-108,264-271; models/jivae.py:199-220; ssivae.py:198-215) */
int pvb_enum_head_fwd(const float* logits, float* alpha, float* w, int64_t B,
                      int K, void* stream);
/* cost[k*B+b] = ll + beta_z*kl (ssiVAE) or ll (jiVAE);
 * elbo_b = sum_k alpha_bk (cost_kb + beta_d (log(1/K) - log alpha_bk));
 * loss_out[0] -= sum_b elbo_b ; glogits = dloss/dlogits through the softmax.
 * glogits must have room for B*K + B floats (the tail is per-sample scratch). */
int pvb_enum_head_bwd(const float* alpha, const float* cost, float beta_d,
                      float* glogits, float* loss_out, int64_t B, int K,
                      void* stream);
/* supervised classification term (ssivae.py:229-242):
 * loss_out[0] -= mult * sum_b log alpha[b, y_b]; glogits written
 * (room for B*K + B floats, as above). */
int pvb_class_nll(const float* logits, const float* y_onehot, float mult,
                  float* glogits, float* loss_out, int64_t B, int K,
                  void* stream);

/* ---- fused small-batch MLP kernels (encoder / classifier side) -----------
 * The guide's encoder is launch-latency bound at SVI batch sizes (M <= a few
 * thousand rows, widths <= 256): these three entry points replace the
 * per-layer GEMM / bias / activation / column-sum launches.
 * Reference: nets/fc.py:51-61 (fcEncoderNet), :97-108 (jfcEncoderNet),
 * :264-271 (fcClassifierNet), models/ivae.py:204-221 (guide) and their autograd. */
#define PVB_MLP_MAX_LAYERS 4
#define PVB_MLP_MAX_HEADS 3
#define PVB_MLP_MAX_WIDTH 256
#define PVB_MLP_MAX_HEAD_DIM 64
typedef struct {
  int64_t M;                 /* rows */
  /* hidden layers computed here, on h_in [M, w_in] (the output of the first,
   * wide layer, computed by pvb_linear_fwd): y_l = act(y_{l-1} W_l^T + b_l) */
  int32_t n_layers;          /* 0..3 */
  int32_t w_in;
  const float* h_in;
  int32_t width[3];
  const float* W[3];
  const float* b[3];
  float* h[3];               /* saved activations [M, width_l] */
  float* pre[3];             /* pre-activations (gelu only; may be NULL) */
  int32_t act;
  /* linear heads on the last hidden activation: out_k = y W_k^T + b_k */
  int32_t n_heads;           /* 1..3 */
  int32_t hdim[3];
  const float* hW[3];
  const float* hb[3];
  float* hout[3];
  /* reparameterised sample from heads 0 (mu) and 1 (s_pre): as pvb_randn + pvb_latent_fwd */
  int32_t gauss;             /* 0/1; Z = hdim[0] = hdim[1] */
  int32_t gen_eps;           /* 1: eps from Philox (seed, *step_counter, first_index); 0: read eps */
  float* eps;                /* [M, Z] read (gen_eps = 0) or written (gen_eps = 1) */
  float* sigma;
  float* z;
  float* kl;
  uint64_t seed;
  const int32_t* step_counter;
  int64_t first_index;
  /* coordinate-transform fold of z into the first decoder layer: as pvb_fold_fwd */
  int32_t fold;              /* 0/1 */
  pvb_fold_cfg cfg;
  const float* cond;
  const float* Wc;
  const float* bc;
  const float* Wz;
  float* Uv;
} pvb_mlp_tail_args;
int pvb_mlp_tail_fwd(const pvb_mlp_tail_args* a, void* stream);

typedef struct {
  int64_t M;
  int32_t n_layers;          /* hidden layers 0..n-1 (1..4), widths width[l] */
  int32_t width[4];
  const float* W[4];         /* W[l] [width[l], width[l-1]]; W[0] unused (no input gradient) */
  const float* h[4];         /* saved activations */
  const float* pre[4];       /* gelu only */
  int32_t act;
  float* dpre[4];            /* out: gradient wrt the pre-activation of layer l, [M, width[l]] */
  int32_t n_heads;
  int32_t hdim[3];
  const float* hW[3];        /* [hdim, width[n-1]] */
  const float* g[3];         /* gradient wrt head output k, [M, hdim[k]] */
} pvb_mlp_chain_args;
/* gradient through the heads and the hidden stack, one launch */
int pvb_mlp_chain_bwd(const pvb_mlp_chain_args* a, void* stream);

typedef struct {
  const float* d;            /* [M, N] gradient wrt the layer output (pre-activation) */
  const float* x;            /* [M, K] layer input */
  float* dW;                 /* [N, K] += d^T x */
  float* db;                 /* [N]    += colsum(d)  (may be NULL) */
  int32_t N, K;
} pvb_wgrad_problem;
/* all weight / bias gradients of a stack in one launch (deterministic: every
 * output element is produced by one CTA in a fixed order) */
int pvb_mlp_wgrad(const pvb_wgrad_problem* problems, int n_problems, int64_t M, void* stream);

/* Per-instance backward of the latent side in one launch: gathers dUv from the
 * fused decoder's per-tile partials (gUv_part; or reads gUv when gUv_part is
 * NULL), pvb_fold_bwd, then pvb_latent_bwd on the resulting dz
 * (gmu / gs_pre written when non-NULL; requires I rows of eps/sigma/s_pre). */
/* part must hold pvb_latent_side_num_partials(I) rows of Hd*(ndim+1+L+C) floats */
int pvb_latent_side_num_partials(int64_t I);
int pvb_latent_side_bwd(const pvb_fold_cfg* cfg, const float* z, const float* cond,
                        const float* Wc, const float* Wz, const float* gUv,
                        const float* gUv_part, int N, float* gz, float* gcond,
                        float* part, const float* eps, const float* sigma,
                        const float* s_pre, const float* w, float beta, float* gmu,
                        float* gs_pre, int64_t I, void* stream);

/* ---- convolutional layers (VED: nets/conv.py:146-249; torch Conv{1,2}d, MaxPool,
 * F.interpolate and their autograd) -------------------------------------------
 * NCHW fp32.  k = 1 or 3, stride 1, zero padding k/2 ("same").  1-D signals
 * [B, C, L]: pass H = 1, kh = 1, W = L, kw = k.  W [Cout, Cin, kh, kw]. */
/* y = act(conv(x, W) + b); pre (optional) receives the pre-activation (gelu) */
int pvb_conv_fwd(const float* x, const float* W, const float* b, float* y,
                 float* pre, int B, int Cin, int Cout, int H, int Wd, int kh,
                 int kw, int act, void* stream);
/* dx = conv_transpose(dpre, W)  (gradient wrt the layer input).  y_below (optional, shape of dx) =
 * OUTPUT of the layer below, act = its activation (not gelu): dx *= act'(y_below), i.e. dx comes
 * out as that layer's dpre. */
int pvb_conv_bwd_data(const float* dpre, const float* W, float* dx, int B,
                      int Cin, int Cout, int H, int Wd, int kh, int kw,
                      const float* y_below, int act, void* stream);
/* dW += dpre (*) x ; db[co] += sum dpre   (accumulate; db may be NULL) */
int pvb_conv_bwd_weight(const float* dpre, const float* x, float* dW, float* db,
                        int B, int Cin, int Cout, int H, int Wd, int kh, int kw,
                        void* stream);
/* dpre = dy * act'(y)  (elementwise; may run in place) */
int pvb_act_bwd(const float* dy, const float* y, const float* pre, float* dpre,
                int64_t n, int act, void* stream);
/* nn.MaxPool{1,2}d(2, 2): x [BC, H, W] -> y [BC, H/2 (2-D) or H, W/2] */
int pvb_maxpool2_fwd(const float* x, float* y, int64_t BC, int H, int Wd,
                     int two_d, void* stream);
/* act != PVB_ACT_NONE: x is the OUTPUT of that activation; dx is then multiplied by the
 * activation derivative at the routed element (= dpre of the layer below; not gelu) */
int pvb_maxpool2_bwd(const float* x, const float* dy, float* dx, int64_t BC,
                     int H, int Wd, int two_d, int act, void* stream);
/* F.interpolate(scale_factor=2): nearest, or bilinear (2-D only, align_corners=False) */
int pvb_upsample2_fwd(const float* x, float* y, int64_t BC, int H, int Wd,
                      int two_d, int bilinear, void* stream);
/* y_below / act as in pvb_conv_bwd_data */
int pvb_upsample2_bwd(const float* dy, float* dx, int64_t BC, int H, int Wd,
                      int two_d, int bilinear, const float* y_below, int act,
                      void* stream);

/* ---- volumetric (3-D) variants of the conv-net layers (csrc/pvb_conv3d.cu; reference
 * nets/conv.py with ndim = 3) ----  NCDHW fp32, cubic kernel k = 1 | 3, stride 1, padding
 * k/2; same contracts as pvb_conv_* / pvb_maxpool2_* / pvb_upsample2_* (nearest only:
 * the reference switches 'bilinear' to 'nearest' for 3-D data, conv.py:127-130). */
int pvb_conv3d_fwd(const float* x, const float* W, const float* b, float* y, float* pre,
                   int B, int Cin, int Cout, int D, int H, int Wd, int k, int act,
                   void* stream);
int pvb_conv3d_bwd_data(const float* dpre, const float* W, float* dx, int B, int Cin,
                        int Cout, int D, int H, int Wd, int k, void* stream);
int pvb_conv3d_bwd_weight(const float* dpre, const float* x, float* dW, float* db, int B,
                          int Cin, int Cout, int D, int H, int Wd, int k, void* stream);
int pvb_maxpool3d_fwd(const float* x, float* y, int64_t BC, int D, int H, int Wd,
                      void* stream);
int pvb_maxpool3d_bwd(const float* x, const float* dy, float* dx, int64_t BC, int D,
                      int H, int Wd, void* stream);
int pvb_upsample3d_fwd(const float* x, float* y, int64_t BC, int D, int H, int Wd,
                       void* stream);
int pvb_upsample3d_bwd(const float* dy, float* dx, int64_t BC, int D, int H, int Wd,
                       void* stream);

/* ---- data-parallel exchange over NVLink peer memory (csrc/pvb_peer.cu) ----
 * Fused SUM all-reduce of the flat [n gradients | loss] buffers of all ranks + the Adam
 * update of pvb_adam_flat_step, one kernel, deterministic (fixed-order sums, identical on
 * every rank).  g: this rank's LOCAL gradient buffer (n + 4 floats: n gradients, the loss
 * accumulator g[n], the last step's global loss g[n+1], 2 pad).  stage_ptrs: DEVICE array of
 * 2 * world pointers, entry parity * world + r = rank r's staging buffer of that parity (n + 4
 * floats each, symmetric memory mapped into this process); peer_flags: DEVICE array of `world`
 * pointers to each rank's flag block of pvb_peer_flag_words() zero-initialised uint32
 * (symmetric memory); state: pvb_peer_state_words() zero-initialised int32 of this rank.
 * two_shot != 0: reduce-scatter + all-gather over peer memory (inbound 2 (world-1)/world x n
 * instead of (world-1) x n floats; pays off for world >= 4); every rank must pass the same value.
 * On return (stream order) p/m/v are updated, g[0..n] is ZERO (ready for the next step's
 * accumulation), g[n+1] holds the global loss, *step_counter is advanced.  There is no
 * end-of-kernel handshake: staging buffers alternate by epoch parity.
 * Every rank must launch it the same number of times (SPMD); n % 4 == 0. */
int pvb_peer_flag_words(void);
int pvb_peer_state_words(void);
int pvb_peer_allreduce_adam(float* p, float* m, float* v, float* g, int64_t n,
                            const void* stage_ptrs, const void* peer_flags, int32_t* state,
                            int rank, int world, int two_shot, float lr, float beta1,
                            float beta2, float eps, int32_t* step_counter,
                            const int32_t* first_step,
                            float* loss_ring /* as pvb_adam_flat_step; may be NULL */,
                            void* stream);

/* ---- nn.BatchNorm{1,2}d of the convolutional nets (csrc/pvb_norm.cu; reference
 * nets/conv.py:187,240 with batchnorm=True, utils/nn.py:103-105) ----------------
 * x, y [B, C, HW] fp32.  training != 0: batch statistics (biased variance) normalise,
 * running_mean / running_var (unbiased) move by `momentum` and *num_batches_tracked
 * (int64; each may be NULL) += 1; training == 0: the running statistics normalise.
 * save_mean / save_invstd [C] receive the statistics used (input of pvb_bn_bwd).
 * workspace: pvb_bn_workspace_bytes(C) bytes, 16-byte aligned. */
int64_t pvb_bn_workspace_bytes(int C);
int pvb_bn_fwd(const float* x, const float* gamma, const float* beta,
               float* running_mean, float* running_var, int64_t* num_batches_tracked,
               float* y, float* save_mean, float* save_invstd, void* workspace, int B,
               int C, int64_t HW, float eps, float momentum, int training, void* stream);
/* backward: dgamma[c] += sum dy xhat, dbeta[c] += sum dy (either may be NULL);
 * training != 0: dx = gamma invstd (dy - mean(dy) - xhat mean(dy xhat));
 * training == 0 (the forward normalised with the running statistics, which are constants):
 * dx = gamma invstd dy.  dx may alias dy */
int pvb_bn_bwd(const float* dy, const float* x, const float* gamma,
               const float* save_mean, const float* save_invstd, float* dx,
               float* dgamma, float* dbeta, void* workspace, int B, int C, int64_t HW,
               int training, void* stream);

/* ---- regression variant (models/ss_reg_ivae.py:172-175,205-207,240-242) ----
 * loss_out[0] += scale * sum_i log N(y_i; loc_i, sigma)   (loc may be NULL = 0;
 * loss_out may be NULL); gloc[i] = scale (y_i - loc_i) / sigma^2 when non-NULL
 * (the gradient of that term wrt loc; scale = -multiplier for a loss term). */
int pvb_normal_logprob(const float* y, const float* loc, float sigma, float scale,
                       float* loss_out, float* gloc, int64_t n, void* stream);
/* dx_cols[M, ncols] (+)= dpre[M, N] W[N, K][:, col0 : col0 + ncols] */
int pvb_linear_dx_cols(const float* dpre, const float* W, float* dx_cols, int64_t M,
                       int N, int K, int col0, int ncols, int accumulate, void* stream);

/* ---- the same convolutions on tcgen05 tensor cores (csrc/pvb_conv_tc.cu) --------
 * For layers with Cin, Cout multiples of 16 (<= 128 outputs per GEMM, <= 256 gathered
 * channels).  Same NCHW fp32 tensors as pvb_conv_*; operands are converted on the fly to
 * fp16, fp32 accumulation in tensor memory.
 * workspace: pvb_conv_tc_workspace_bytes(...) bytes, 16-byte aligned (repacked weights). */
int pvb_conv_tc_supported(int Cin, int Cout, int kh, int kw);
int pvb_conv_tc_wgrad_supported(int Cin, int Cout, int kh, int kw);  /* also Cin < 16 (padded) */
int64_t pvb_conv_tc_workspace_bytes(int Cin, int Cout, int kh, int kw);
/* mode 0: dst = act(conv(src = x, W) + b), pre optional; mode 1: dst = dx from src = dpre.
 * mode | 2: the workspace already holds this mode's repacked weights (pvb_conv_tc_prep, e.g. once
 * per optimizer step off the critical path); otherwise the call repacks them first. */
int pvb_conv_tc_pix(const float* src, const float* W, const float* b, float* dst,
                    float* pre, void* workspace, int B, int Cin, int Cout, int H,
                    int Wd, int kh, int kw, int act, int mode, void* stream);
int pvb_conv_tc_prep(const float* W, void* workspace, int Cin, int Cout, int kh, int kw,
                     int mode, void* stream);
/* dW += dpre (*) x ; db += sum dpre  (atomic accumulation).
 * scratch (optional): pvb_conv_tc_wgrad_scratch_bytes(...) bytes of ZEROED device memory, left
 * zeroed on return (reusable by the next call on the same stream, for any layer that fits).  With
 * it the per-CTA accumulators are added coalesced into a transposed copy and folded into dW by a
 * small kernel (below); without it they are added straight into dW (strided, slower). */
int64_t pvb_conv_tc_wgrad_scratch_bytes(int Cin, int Cout, int kh, int kw);
/* fold != 0: the scratch copy is folded into dW / db (and cleared) before returning; fold == 0: the
 * sums stay in this layer's OWN scratch until pvb_conv_tc_wgrad_fold, which folds up to any number
 * of layers in one launch (a backward pass: one fold for all its layers). */
int pvb_conv_tc_wgrad(const float* dpre, const float* x, float* dW, float* db, int B,
                      int Cin, int Cout, int H, int Wd, int kh, int kw, void* scratch,
                      int fold, void* stream);
#define PVB_WGRAD_FOLD_MAX 16      /* layers per launch (more are folded by further launches) */
typedef struct {
  float* scratch;                  /* the layer's scratch (pvb_conv_tc_wgrad_scratch_bytes) */
  float* dW;                       /* [Cout][Cin][taps], accumulated into */
  float* db;                       /* [Cout] or NULL */
  int Cin, Cout, taps;
} pvb_wgrad_fold;
int pvb_conv_tc_wgrad_fold(const pvb_wgrad_fold* layers, int n, void* stream);

/* ---- optimizer / reductions -------------------------------------------- */
/* out[j] (+)= sum_g part[g*part_stride + j], j < n, fixed order (deterministic) */
int pvb_reduce_partials(const float* part, float* out, int G, int64_t n,
                        int64_t part_stride, int accumulate, void* stream);
int pvb_counter_add(int32_t* counter, int32_t v, void* stream);
/* torch.optim.Adam defaults (Pyro optim.Adam, trainers/svi.py:79-81).
 * Pyro keeps one optimizer PER PARAMETER, created when the parameter first
 * carries a gradient: element i takes its own step count
 *   t_i = *step_counter - first_step[i]   (step_counter already incremented),
 * and is left untouched while first_step[i] < 0 (parameter not yet seen).
 * first_step == NULL: every element active since step 0. */
int pvb_adam_flat(float* p, const float* g, float* m, float* v, int64_t n,
                  float lr, float beta1, float beta2, float eps,
                  const int32_t* step_counter, const int32_t* first_step,
                  void* stream);
/* Same update with the step increment folded in: every element uses
 * t = *step_counter + 1 (- first_step[i]); the last CTA to finish stores
 * *step_counter += 1 (ticket must point to a zeroed int32 owned by the caller).
 * g[0..n) is ZEROED as it is consumed (Pyro zeroes the gradients after every optimizer step,
 * svi.step; here it also saves the next step's memset).
 * loss_src (optional): two floats {loss accumulator of this step, last loss}: the accumulator
 * is copied to loss_src[1] and cleared.  loss_ring (optional, needs loss_src): PVB_LOSS_RING
 * floats of device-accessible memory, normally MAPPED PINNED HOST memory; slot (new step count &
 * (PVB_LOSS_RING - 1)) receives the step's loss, i.e. the result reaches the host without a
 * separate copy. */
#define PVB_LOSS_RING 16
int pvb_adam_flat_step(float* p, float* g, float* m, float* v, int64_t n,
                       float lr, float beta1, float beta2, float eps,
                       int32_t* step_counter, const int32_t* first_step,
                       int32_t* ticket, float* loss_src, float* loss_ring,
                       void* stream);

/* dst[r][:] = src[idx[r]][:] for r < rows (row_floats fp32 each; 16-byte accesses when
 * row_floats % 4 == 0 and both buffers are 16-byte aligned): the on-device shuffle of the
 * GPU-resident batch loader (replaces the DataLoader's sampler + collate of the reference,
 * utils/data.py:6-52), HBM-bound: 8 B per element. idx: int64, values in [0, n_src). */
int pvb_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t rows,
                    int64_t row_floats, int64_t n_src, void* stream);

/* ---- spatial decoder, fused tcgen05 path (Hd = 128, two tanh layers) ----
 * One persistent kernel per step: grid -> h0 -> (128x128 tcgen05 GEMM + tanh)
 * x2 -> out layer -> log-lik -> full backward, activations never leave the
 * SM (nets/fc.py:189-237 forward; autograd backward of the same).
 * Outputs: rowll[R], loc[R] (optional), and when `backward` != 0:
 *   gUv_part [T][slots=5][3][128]  per-tile partial sums (T = #tiles of 128 rows)
 *   wgrad_part [G][PVB_TC_WGRAD_STRIDE]  per-CTA weight-gradient partials, laid out
 *              (dW1[128][128] | db1 | dW2[128][128] | db2 | dwo[128] | dbo | pad)
 * Workspace sizes via pvb_sdec_tc_sizes. */
typedef struct {
  int64_t tiles;         /* T */
  int32_t ctas;          /* G */
  int64_t gUv_part_floats;
  int64_t wgrad_part_floats; /* G * PVB_TC_WGRAD_STRIDE */
} pvb_tc_sizes;
#define PVB_TC_TILE 128       /* rows per tile of the fused kernel */
#define PVB_TC_MAX_SLOTS 5    /* instances a tile can touch (N >= 32) */
#define PVB_TC_WGRAD_FLOATS (2 * 128 * 128 + 2 * 128 + 128 + 1)
/* per-CTA stride of wgrad_part (16-byte aligned rows) */
#define PVB_TC_WGRAD_STRIDE ((PVB_TC_WGRAD_FLOATS + 3) / 4 * 4)
int pvb_sdec_tc_sizes(int64_t I, int N, pvb_tc_sizes* out);
int pvb_sdec_tc_step(const float* Uv, const float* x, const float* w,
                     const float* W1, const float* b1, const float* W2,
                     const float* b2, const float* wo, const float* bo,
                     float* rowll, float* loc, float* gUv_part,
                     float* wgrad_part, int64_t I, int64_t B, int H, int W,
                     int ndim, int sampler, int sigmoid_d, float decoder_sig,
                     int backward,
                     const void* packed_w /* optional: pvb_sdec_tc_pack_weights output; then the
                                             weight tiles come in by two TMA bulk copies
                                             (cp.async.bulk) instead of being converted by
                                             every CTA; NULL: converted from W1 / W2 */,
                     void* stream);
/* W1, W2 (fp32 [128][128]) -> the kernel's fp16 operand tiles, pvb_sdec_tc_packed_weight_bytes()
 * bytes (W1 tile, then W2 tile); run once per optimizer step, off the critical path. */
int64_t pvb_sdec_tc_packed_weight_bytes(void);
int pvb_sdec_tc_pack_weights(const float* W1, const float* W2, void* packed, void* stream);
/* gUv_part -> gUv [I,3,128] (deterministic) */
int pvb_sdec_tc_gather_gUv(const float* gUv_part, float* gUv, int64_t I, int N,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVB_H_ */
