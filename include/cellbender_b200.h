/* cellbender_b200 -- C ABI of the B200-native (sm_100a) hot path of CellBender `remove-background`.
 *
 * The reference (broadinstitute/CellBender, pure Python/PyTorch/Pyro) has no FFI: its boundary is a set of
 * Python call sites.  Every entry point below names the reference call site it
 * replaces (paths relative to /root/reference/cellbender/remove_background/).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add at each site.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types.  Every pointer is a DEVICE pointer unless its
 *     name ends in _h.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all calls are asynchronous on `stream`; the caller owns every buffer (allocate through any allocator).
 *   - return value: 0 = launched OK; non-zero = error, message in cb_last_error() (thread-local).
 *   - dense matrices are row-major fp32 with a leading dimension `ld` that is a multiple of 4 (16-byte rows).
 *   - CSR count matrices: indptr int64[R+1], indices int32[nnz] (sorted, unique per row), data fp32[nnz].
 *   - "m" indices (flattened count-matrix index n * G_all + g) are int64 (posterior.py:1554-1575).
 */
#ifndef CELLBENDER_B200_H_
#define CELLBENDER_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

int cb_abi_version(void);
const char* cb_last_error(void);

/* ---------------------------------------------------------------------------------------------------
 * A1/A2  minibatch loader.  Replaces DataLoader.__next__ + sparse_collate (data/dataprep.py:153-193,
 * 275-292), transform_input (vae/encoder.py:343-381) and the count features of
 * EncodeNonZLatents.forward (vae/encoder.py:250-272).
 *   row_ids[n_rows]   rows of the device-resident CSR chosen by the host RNG (np.random order preserved)
 *   amb_unit[G]       unit-L2-norm ambient profile (encoder.py:266-270), may be NULL
 *   x_raw/x_norm/x_lognorm [n_rows, ld]   counts, x/(sum+1e-5), log1p(x/(sum+1e-5)); any may be NULL
 *   row_feats[n_rows, 4]                  sum(x), #(x>0), dot(x, amb_unit), max(x); may be NULL          */
int cb_csr_gather_rows(const long long* indptr, const int* indices, const float* data, const long long* row_ids,
                       int n_rows, int n_genes, long long ld, const float* amb_unit, float* x_raw, float* x_norm,
                       float* x_lognorm, float* row_feats, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A7  NegativeBinomialPoissonConvApprox.log_prob(value)  (distributions/NegativeBinomialPoissonConvApprox.py:105-143)
 * and its backward.  Logical shape [E, B, G]; s* are element strides {e, b, g} of each operand (0 = broadcast),
 * i.e. value [B,G], mu/lam [2,B,G] and a scalar alpha as pyro's enumeration produces them (model.py:338-346).
 * nan_flag: int32 on device, OR-ed with 1 if any output is NaN (the host raises NanException, ...Approx.py:129-141). */
int cb_nbpc_approx_log_prob_fwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                                const float* alpha, const long long* sa, const float* lam, const long long* sl,
                                float* out, int E, int B, int G, int* nan_flag, void* stream);
int cb_nbpc_approx_log_prob_bwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                                const float* alpha, const long long* sa, const float* lam, const long long* sl,
                                const float* grad_out, float* dmu, float* dlam, float* dalpha, int E, int B, int G,
                                void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A8  NegativeBinomialPoissonConv.log_prob(value)  (distributions/NegativeBinomialPoissonConv.py:89-154): the exact
 * NB (+) Poisson convolution, logsumexp over max_poisson + 1 (reference: 101) Poisson counts starting at
 * clamp(min(int(lam - max_poisson/2), int(value - max_poisson)), 0); used when consts.USE_EXACT_LOG_PROB is set.
 * Same operand conventions as cb_nbpc_approx_log_prob_fwd/_bwd.                                                */
int cb_nbpc_exact_log_prob_fwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                               const float* alpha, const long long* sa, const float* lam, const long long* sl,
                               float* out, int E, int B, int G, int max_poisson, int* nan_flag, void* stream);
int cb_nbpc_exact_log_prob_bwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                               const float* alpha, const long long* sa, const float* lam, const long long* sl,
                               const float* grad_out, float* dmu, float* dlam, float* dalpha, int E, int B, int G,
                               int max_poisson, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A5(softmax)+A6+A7+K7  observation terms of the ELBO, forward and backward fused.  Replaces, per svi.step:
 * Decoder softmax (vae/decoder.py:46-47), calculate_mu/calculate_lambda with y enumerated (model.py:25-85,308-320),
 * NBPCapprox(...).to_event(1).log_prob(x) (model.py:338-346), the Poisson "obs_empty" site (model.py:374-398) and
 * their autograd backward.
 *   coef[n_rows, 7]   a_mu1, cA1, cB1, cA0, cB0, cAE, cBE:  mu_y1 = a_mu1*chi + 1e-10, mu_y0 = 1e-10,
 *                     lam_* = cA* * chi_ambient + cB* * chi_bar + 1e-10
 *   alpha[1]          1/phi + 1e-10 (model.py:334,341)
 *   w[n_rows, 3]      d loss / d (L_y1, L_y0, L_E)
 *   sums[n_rows, 12]  L_y1, L_y0, L_E, dL1/da_mu1, dL1/dcA1, dL1/dcB1, dL1/dalpha, dL0/dcA0, dL0/dcB0, dL0/dalpha,
 *                     dLE/dcAE, dLE/dcBE
 *   dlogits[n_rows, ldl]  d loss / d logits;   qamb[n_rows, ldq]  per-row d loss / d chi_ambient (sum rows with
 *                     cb_colsum_rows);   lse_out[n_rows] row logsumexp of logits (may be NULL)
 *   scratch[cb_elbo_obs_scratch_elems(n_rows)]  floats, 8-byte aligned: per (gene range, droplet) tile the partial row
 *                     sums in float64 and the (max, sum exp) pair of its logits -- a droplet row is split into gene
 *                     ranges and the three phases (row logsumexp; likelihood + gradients; softmax backward) are three
 *                     launches over the tiles, ordered by the stream.                                              */
long long cb_elbo_obs_scratch_elems(int n_rows);
int cb_elbo_obs_fwd_bwd(const long long* indptr, const int* indices, const float* data, const long long* row_ids,
                        int n_rows, int n_genes, const float* logits, long long ldl, const float* chi_ambient,
                        const float* chi_bar, const float* coef, const float* alpha, const float* w, float* sums,
                        float* dlogits, float* qamb, long long ldq, float* lse_out, int* nan_flag, float* scratch,
                        void* stream);
/* The same with the exact NB (+) Poisson convolution as the likelihood: NegativeBinomialPoissonConv(max_poisson)
 * (distributions/NegativeBinomialPoissonConv.py:89-154; model.py:336-357 with consts.USE_EXACT_LOG_PROB = True).  Genes
 * without a stored count use the one-term closed form, stored counts the term-by-term sum in float64.               */
int cb_elbo_obs_exact_fwd_bwd(const long long* indptr, const int* indices, const float* data, const long long* row_ids,
                        int n_rows, int n_genes, const float* logits, long long ldl, const float* chi_ambient,
                        const float* chi_bar, const float* coef, const float* alpha, const float* w, float* sums,
                        float* dlogits, float* qamb, long long ldq, float* lse_out, int* nan_flag, float* scratch,
                        int max_poisson, void* stream);
int cb_colsum_rows(const float* in, int n_rows, int n_cols, long long ld, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A4  self.dropout50(self.batchnorm0(x)) of EncodeNonZLatents.forward (vae/encoder.py:239,279), forward + backward.
 * x / y / dy are [n_rows, ld] fp32; keep[n_rows] is the Dropout1d row realisation (0 or 2; NULL = none);
 * running_mean/var are updated in training mode (momentum 0.1, unbiased variance, as nn.BatchNorm1d) and
 * *num_batches_tracked (may be NULL) is incremented.                                                             */
int cb_bn0_forward(const float* x, long long ldx, int n_rows, int n_genes, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, long long* num_batches_tracked, const float* keep, int training,
                   float momentum, float eps, float* y, long long ldy, float* save_mean, float* save_rstd, void* stream);
int cb_bn0_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_genes,
                    const float* keep, const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A4/A5  nn.BatchNorm1d + activation of the hidden layers ([n_rows, n_cols = 512]): EncodeNonZLatents.batchnorm1/2 +
 * softplus (vae/encoder.py:283-289), eps_network BatchNorm1d + Softplus (vae/encoder.py:219-228), the Decoder's
 * FullyConnectedLayer BatchNorm1d(momentum 0.01, eps 1e-3) + ReLU (vae/base.py:66-79); forward and backward.
 * act: 0 identity, 1 softplus, 2 relu.  training = 1: minibatch statistics, running stats and *num_batches_tracked
 * (may be NULL) updated.  backward: dx [n_rows, lddx], dgamma, dbeta [n_cols]; dbias (may be NULL) = column sums of dx,
 * the gradient of the Linear bias that fed the BatchNorm.
 * feat[n_rows, n_feat] / feat_weight[n_cols, ldw] (may be NULL): the hand-made feature columns of Linear(cat(feat, x_in))
 * (vae/encoder.py:283-298) enter here, x[r, c] += sum_f feat[r, f] * feat_weight[c, f] (written back to x), so that the
 * gene-wide product that produced x need not wait for the features.                                             */
int cb_bn_act_forward(float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long long* num_batches_tracked, int training,
                      float momentum, float eps, int act, float* y, long long ldy, float* save_mean, float* save_rstd,
                      const float* feat, int n_feat, const float* feat_weight, long long ldw, void* stream);
int cb_bn_act_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_cols,
                       const float* gamma, const float* beta, const float* save_mean, const float* save_rstd, int training,
                       int act, float* dx, long long lddx, float* dgamma, float* dbeta, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A4  small per-droplet pieces of EncodeNonZLatents.forward (vae/encoder.py:250-293) and of the observation term.
 * cb_encoder_features: feats[n_rows, 4] (cb_csr_gather_rows), z[n_rows, ldz] (EncodeZ loc) ->
 *     p_feat[n_rows, 4] = log1p(sum), log1p(nnz), -0.5 (max(log1p(sum) - d_empty_loc, 0) / 0.1)^2, ||z||_2
 *     e_feat[n_rows, 4] = log1p(sum), log1p(nnz), dot(x, ambient unit vector), ||z||_2
 *     d_empty_loc[1]: CONSTRAINED pyro.param value (encoder.py:262); NULL = no ambient profile (features 2 are 0).
 * cb_head_forward / _backward: out[r] = bias + <w, cat(feat[r, :n_feat], x[r, :n_cols])>, the Linear(.. -> 1) heads
 *     layer3 (encoder.py:216) and eps_network[-1] (encoder.py:227); backward: dx[n_rows, lddx] (may be NULL),
 *     dw[n_feat + n_cols], dbias[1] (may be NULL).
 * cb_obs_loss: loss[1] = sum_b w[b, :] . sums[b, 0:3], neg_loss[1] = -loss (may be NULL)  (sums of cb_elbo_obs_fwd_bwd).
 * cb_obs_small_grads: g_up[1] (NULL = 1) upstream gradient -> d_coef[n_rows, 7], d_alpha[1], d_w[n_rows, 3] (may be NULL). */
int cb_encoder_features(const float* feats, const float* z, long long ldz, int n_rows, int z_dim, const float* d_empty_loc,
                        float* p_feat, float* e_feat, void* stream);
int cb_head_forward(const float* x, long long ldx, const float* feat, int n_feat, int n_rows, int n_cols, const float* w,
                    const float* bias, float* out, void* stream);
int cb_head_backward(const float* x, long long ldx, const float* feat, int n_feat, int n_rows, int n_cols, const float* w,
                     const float* d_out, float* dx, long long lddx, float* dw, float* dbias, void* stream);
int cb_obs_loss(const float* sums, int ld_sums, const float* w, int n_rows, float* loss, float* neg_loss, void* stream);
int cb_obs_small_grads(const float* sums, int ld_sums, const float* w, const float* g_up, int n_rows, float* d_coef,
                       float* d_alpha, float* d_w, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A3-A5  the large contractions of EncodeZ / EncodeNonZLatents / Decoder (vae/encoder.py:97-125, 211-228, 279-298;
 * vae/decoder.py:41-47) and their backward, on tcgen05 tensor cores (TF32 multiply, fp32 accumulate in TMEM).
 *   C[M, N] (=, +=) A[M, K] * B[N, K]^T (+ bias[N])
 * a_mn / b_mn = 0: operand stored [M|N rows, K cols] (K-major, the nn.Linear forward case);
 *             = 1: operand stored [K rows, M|N cols] (what dW = dY^T X and dX = dY W see, no transposes in memory).
 * lda / ldb / ldc multiples of 4 and 16-byte aligned bases (TMA loads and TMA stores); out-of-range rows/cols are
 * zero-filled on load and clipped on store (rows exactly, columns at 16-byte granularity: up to 3 padding columns behind
 * column N - 1 of a row -- inside the row's padded pitch -- may be overwritten with zeros); accumulate uses the TMA
 * unit's element-wise reduce-add.
 * split_k > 1 streams K across CTAs (skinny M): partial tiles in `workspace`
 * (cb_gemm_tf32_workspace_elems floats), summed in a fixed order.                                            */
long long cb_gemm_tf32_workspace_elems(int M, int N, int split_k);
int cb_gemm_tf32_suggest_split(int M, int N, int K);
/* C = A B^T + A2 B2^T (two operand pairs share one TMEM accumulator; K2 = 0 disables the second pair): the input
 * gradient of two layers that read the same input, dx = dy1 W1 + dy2 W2, in one launch with no read-modify-write. */
int cb_gemm_tf32_pair(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                      long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                      long long ldc, int M, int N, const float* bias, int accumulate, int split_k, float* workspace,
                      long long workspace_elems, void* stream);
int cb_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* C,
                 long long ldc, int M, int N, int K, const float* bias, int accumulate, int split_k, float* workspace,
                 long long workspace_elems, void* stream);
/* The same product with an explicit tile configuration (tuning / benchmarking): tile_cfg 0 = heuristic, 1 = 256 x 128
 * tile, 2-stage ring, two CTAs per SM; 2 = 256 x 128 deep ring; 3 = 256 x 256 deep ring; 4 / 5 = 128 x 256 (2-stage / deep);
 * 6 / 7 = 128 x 128 (2-stage / deep).  K2 = 0: single operand pair.                                                */
int cb_gemm_tf32_ex(const float* A, long long lda, const float* B, long long ldb, int K, const float* A2,
                    long long lda2, const float* B2, long long ldb2, int K2, int a_mn, int b_mn, float* C,
                    long long ldc, int M, int N, const float* bias, int accumulate, int split_k, float* workspace,
                    long long workspace_elems, int tile_cfg, void* stream);
/* Weight gradient of the hand-made feature columns of Linear(cat(feat, x)) (vae/encoder.py:283-298, autograd of
 * layer1 / layer2 / eps_network.0): dw[h, f] = sum_b dy[b, h] feat[b, f] into the first n_feat columns of dw [n_out, ldw]. */
int cb_feat_dw(const float* dy, long long ldy, const float* feat, int n_feat, int n_rows, int n_out, float* dw,
               long long ldw, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A4(tail)+A9  the per-droplet latent algebra of RemoveBackgroundPyroModel.model / .guide (model.py:232-445, 447-574),
 * the output shaping of EncodeNonZLatents.forward (vae/encoder.py:300-340), the constraint transforms of the scalar
 * pyro.param sites (model.py:468-513) and the autograd backward of all of it, as one-CTA kernels.
 *   feats[n_rows, 4]        row features of cb_csr_gather_rows (col 0 = total counts)
 *   p_out/eps_out[n_rows]   raw outputs of EncodeNonZLatents.layer3 / eps_network (before offset and gating)
 *   z_loc/z_log_scale[n_rows, ldz]  EncodeZ.loc_out / sig_out outputs (scale = exp(log_scale), encoder.py:121-124)
 *   u_*                     UNCONSTRAINED scalar parameters (device, 1 element each; u_rho_* NULL for model 'ambient')
 *   consts_h[cb_latent_num_consts()]  HOST array: log_count_crossover, prior_log_cell_counts, empty_log_count_threshold,
 *                           offset_logit_p, d_cell_loc_prior, d_cell_scale_prior, epsilon_prior, phi_conc_prior,
 *                           phi_rate_prior, rho_alpha_prior, rho_beta_prior, d_empty_loc_prior, d_empty_scale_prior,
 *                           p_logit_prior, empty_UMI_threshold, counts_crossover, rho_param_min, rho_param_max,
 *                           model_type (0 ambient, 1 swapping, 2 full)
 * cb_latent_prepare  -> conc_gamma[1 + 3 n_rows] = (phi conc, epsilon conc[n_rows], rho_alpha x n_rows, rho_beta x n_rows):
 *                       the concentrations at which the caller draws standard-gamma samples (Gamma.rsample / Beta.rsample);
 *                       conc_beta/total_beta[n_rows, 2] = (alpha, beta) and alpha + beta, operands of the Dirichlet
 *                       implicit gradient.
 * cb_latent_beta_sample  clamps the draws to >= FLT_MIN and forms xx[n_rows, 2] = (rho, 1 - rho), rho = ga / (ga + gb).
 * cb_latent_forward  noise_z[n_rows, z_dim], noise_row[n_rows, 3] (d_cell, d_empty, p_logit_reg) standard normals;
 *                       gamma_draws[1 + n_rows] (phi, epsilon) ->
 *                       z_out[n_rows, z_dim], coef[n_rows, 7] / alpha[1] / w[n_rows, 3] (operands of cb_elbo_obs_fwd_bwd),
 *                       terms[cb_latent_num_terms()] = phi, z, d_cell, rho, epsilon, d_empty, y, p_logit_reg,
 *                       obs_probably_empty_y, obs_probably_cell_y, epsilon_mean, epsilon_empty_mean; loss_local[1] = -(their sum);
 *                       enc_out[n_rows, 4] = p_y, d_loc, epsilon (encoder), epsilon (sample); med_idx[2] median rows.
 * cb_latent_backward gamma_grad[1 + n_rows] = torch._standard_gamma_grad(conc, draw); dirichlet_grad[n_rows, 2] =
 *                       torch._dirichlet_grad(xx, conc_beta, total_beta); d_z/d_coef/d_alpha/d_w = d loss / d (outputs)
 *                       (any may be NULL = 0); g_local[1] = d loss / d loss_local (NULL = 1) ->
 *                       d_p_out, d_eps_out [n_rows]; d_z_loc, d_z_log_scale [n_rows, lddz]; g_* = d loss / d u_* (NULL ok). */
int cb_latent_num_consts(void);
int cb_latent_num_terms(void);
int cb_latent_prepare(const float* feats, const float* p_out, const float* eps_out, int n_rows,
                      const float* u_d_cell_scale, const float* u_d_empty_loc, const float* u_d_empty_scale,
                      const float* u_rho_alpha, const float* u_rho_beta, const float* u_phi_loc, const float* u_phi_scale,
                      const float* consts_h, float* conc_gamma, float* conc_beta, float* total_beta, void* stream);
int cb_latent_beta_sample(float* gamma_draws, int n_rows, float* xx, void* stream);
int cb_latent_forward(const float* feats, const float* p_out, const float* eps_out, const float* z_loc,
                      const float* z_log_scale, long long ldz, int n_rows, int z_dim, const float* noise_z,
                      const float* noise_row, const float* gamma_draws, const float* xx, const float* u_d_cell_scale,
                      const float* u_d_empty_loc, const float* u_d_empty_scale, const float* u_rho_alpha,
                      const float* u_rho_beta, const float* u_phi_loc, const float* u_phi_scale, const float* consts_h,
                      float* z_out, float* coef, float* alpha, float* w, float* terms, float* loss_local, float* enc_out,
                      int* med_idx, void* stream);
int cb_latent_backward(const float* feats, const float* p_out, const float* eps_out, const float* z_loc,
                       const float* z_log_scale, long long ldz, int n_rows, int z_dim, const float* noise_z,
                       const float* noise_row, const float* gamma_draws, const float* xx, const float* gamma_grad,
                       const float* dirichlet_grad, const int* med_idx, const float* u_d_cell_scale,
                       const float* u_d_empty_loc, const float* u_d_empty_scale, const float* u_rho_alpha,
                       const float* u_rho_beta, const float* u_phi_loc, const float* u_phi_scale, const float* consts_h,
                       const float* d_z, const float* d_coef, const float* d_alpha, const float* d_w, const float* g_local,
                       float* d_p_out, float* d_eps_out, float* d_z_loc, float* d_z_log_scale, long long lddz,
                       float* g_d_cell_scale, float* g_d_empty_loc, float* g_d_empty_scale, float* g_rho_alpha,
                       float* g_rho_beta, float* g_phi_loc, float* g_phi_scale, void* stream);
/* transform_to(constraints.simplex) of the chi_ambient parameter (model.py:245-249, 486-490): torch's
 * SoftmaxTransform.  u[n] -> y[n] = softmax(u) (+ y_unit = y / ||y||_2, the encoder's ambient unit vector,
 * vae/encoder.py:266-270; may be NULL); backward: grad_u = y * (grad_y - <grad_y, y>).                          */
int cb_simplex_forward(const float* u, int n, float* y, float* y_unit, void* stream);
int cb_simplex_backward(int n, const float* y, const float* grad_y, float* grad_u, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K8  fused ClippedAdam + OneCycleLR over one flat buffer.  Replaces get_optimizer (run.py:593-629) and the
 * per-parameter optimizer/scheduler steps at train.py:56-60.  lr_table/beta1_table[table_len] hold torch
 * OneCycleLR's per-step values, step_size_table[table_len] (may be NULL: evaluated on the device) the bias-corrected
 * step size lr sqrt(1 - beta2^t) / (1 - beta1^t) of step t; *step_ptr (device int64) counts completed steps and is
 * incremented here.                                                                                             */
int cb_clipped_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                         const float* beta1_table, const float* step_size_table, long long table_len, long long* step_ptr,
                         float beta2, float eps, float clip, float grad_scale, void* stream);
/* the same update on one contiguous range (step counter advanced only if advance != 0; n = 0 allowed): the sweep of one
 * large weight block right behind its weight-gradient product, and of the ranges between such blocks.              */
int cb_clipped_adam_range(float* p, const float* g, float* m, float* v, long long n, const float* lr_table,
                          const float* beta1_table, const float* step_size_table, long long table_len, long long* step_ptr,
                          float beta2, float eps, float clip, int advance, void* stream);
/* Data-parallel exchange fused into the optimizer sweep (new surface, the reference is single-GPU): ClippedAdam on
 * this rank's contiguous shard, reading the shard's gradient from every rank's gradient buffer (peer memory over
 * NVLink, summed in rank order) and storing the updated parameters into every rank's parameter buffer.
 * grad_ptrs[n_grad] / peer_param_ptrs[n_peer]: HOST arrays of device (peer-mapped) pointers; grad_ptrs[own] is this
 * rank's own buffer (read from HBM, the others are fetched with TMA bulk copies over NVLink).  n % 4 == 0.          */
int cb_clipped_adam_p2p(float* p, float* m, float* v, long long n, const float* const* grad_ptrs, int n_grad, int own,
                        float* const* peer_param_ptrs, int n_peer, const float* lr_table, const float* beta1_table,
                        long long table_len, long long* step_ptr, float beta2, float eps, float clip, int advance,
                        void* stream);
/* The same exchange through the NVSwitch (NVLS): mc_grad / mc_param = multicast addresses of the shard in the symmetric
 * gradient / parameter buffers; multimem.ld_reduce sums the gradients inside the switch, multimem.st delivers the new
 * parameters to every rank.                                                                                         */
int cb_clipped_adam_nvls(float* p, float* m, float* v, long long n, const float* mc_grad, float* mc_param,
                         const float* lr_table, const float* beta1_table, long long table_len, long long* step_ptr,
                         float beta2, float eps, float clip, int advance, void* stream);
/* The same with the transfer paths chosen explicitly (tools/nvls_microbench.py measures them; cb_clipped_adam_nvls is
 * ld_mode 0, st_mode 0, 6 CTAs per SM).  ld_mode: 0 = multimem.ld_reduce, 1 = local_grad only (diagnostic).  st_mode:
 * 0 = multimem.st, 1 = local store + posted P2P stores to peer_param_ptrs[n_peer], 2 = local store only (diagnostic). */
int cb_clipped_adam_nvls_ex(float* p, float* m, float* v, long long n, const float* mc_grad, float* mc_param,
                            const float* local_grad, float* const* peer_param_ptrs, int n_peer, int ld_mode, int st_mode,
                            int blocks_per_sm, const float* lr_table, const float* beta1_table, long long table_len,
                            long long* step_ptr, float beta2, float eps, float clip, int advance, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A10-A12  posterior noise-count log-probability over the stored (non-zero) counts.  Replaces
 * Posterior._log_prob_noise_count_tensor (posterior.py:784-859), the normalisation and running log-mean of
 * noise_log_pdf (posterior.py:698-735) and the thresholding/sparsify/index translation of
 * _get_cell_noise_count_posterior_coo (posterior.py:528-576), dense_to_sparse_op_torch (sparse_utils.py:10-33),
 * IndexConverter.get_m_indices (posterior.py:1554-1575).
 *   cell_rows[n_cells]     CSR rows of the cells (p > 0.5) of this chunk
 *   logits_t[n_genes, ldt] decoder output of the chunk, TRANSPOSED and cell-major: logits_t[g, b * n_samples + s]
 *                          (cb_gemm_tf32 with A = decoder weight, B = hidden activations; bias NOT included);
 *   dec_bias[n_genes]      decoder output bias (may be NULL);  lse[n_cells * n_samples] = cb_col_logsumexp(logits_t, dec_bias)
 *   a_mu/c_a/c_b [n_cells * n_samples] (index b * n_samples + s), alpha[n_samples]:
 *                          mu = a_mu * chi, lam = c_a * chi_ambient + c_b * chi_bar
 *   n_window[n_cells]      min(n_counts_max, max count of the reference minibatch of the cell) (posterior.py:816)
 *   entry_start[n_cells+1] prefix sums of the cells' stored-count numbers (device); entry_start[n_cells] entries are
 *                          processed, max_entries (host) only sizes the grid and may be an upper bound, so that one
 *                          captured CUDA graph serves every chunk of a pass
 *   out_offset             device scalar (may be NULL = 0): first row of this chunk inside the output arrays
 *   barcode_all[n_cells] / gene_all[n_genes]  indices into the full count matrix (IndexConverter, posterior.py:1554-1575)
 * outputs, one row per stored count in (cell, gene) order:
 *   logp_out[n_entries, stride] (stride = cb_posterior_window_stride()), low_out, keep_cnt, entry_m[n_entries]
 * cb_posterior_offset_flags: flag[i] = keep_cnt[i] > 0 && low[i] != 0 (entries that carry a noise-count offset).
 * cb_posterior_compact: thresholded COO triplets + (m, offset) pairs from exclusive scans of keep_cnt / flag.       *   ctas_per_sm            0: one CTA per 64 stored counts; k > 0: at most k CTAs per SM, tiles strided over them (the
 *                          kernel then shares the GPU with the decoder product of the next chunk)
 */
int cb_posterior_window_stride(int n_counts_max);
int cb_posterior_noise_window(const long long* indptr, const int* indices, const float* data, const long long* cell_rows,
                              int n_cells, int n_genes, int n_samples, const float* logits_t, long long ldt,
                              const float* dec_bias, const float* lse, const float* chi_ambient, const float* chi_bar,
                              const float* a_mu, const float* c_a, const float* c_b, const float* alpha,
                              const int* n_window, const long long* entry_start, long long max_entries,
                              const long long* barcode_all, const long long* gene_all, long long total_n_genes,
                              int n_counts_max, float log_thresh, float lambda_multiplier, const long long* out_offset,
                              float* logp_out, int* low_out, int* keep_cnt, long long* entry_m, int ctas_per_sm,
                              void* stream);
int cb_posterior_offset_flags(long long n_entries, const int* keep_cnt, const int* low, int* flag, void* stream);
/* out[c] = log sum_r exp(in[r, c] + row_add[r]) over an [n_rows, ld] matrix (row_add may be NULL): the softmax
 * denominators of Decoder.forward (vae/decoder.py:46-47) for the transposed logits of a posterior chunk.
 * scratch: cb_col_logsumexp_scratch_elems(n_rows, n_cols) floats.                                                   */
long long cb_col_logsumexp_scratch_elems(int n_rows, int n_cols);
int cb_col_logsumexp(const float* in, int n_rows, int n_cols, long long ld, const float* row_add, float* scratch,
                     float* out, void* stream);
int cb_posterior_compact(long long n_entries, int n_counts_max, const long long* entry_m, const float* logp,
                         const int* low, const int* keep_cnt, const long long* keep_pos, const long long* off_pos,
                         float log_thresh, long long* coo_m, int* coo_c, float* coo_lp, long long* off_m, int* off_v,
                         void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A13-A15  estimators over the (m, c)-sorted posterior COO.  Replace apply_function_dense_chunks
 * (estimation.py:747-793), MAP/Mean/ThresholdCDF/SingleSample (estimation.py:89-204),
 * _estimation_array_to_csr (estimation.py:207-235) and MultipleChoiceKnapsack._chunk_estimate_noise
 * (estimation.py:552-702).  mode: 0 MAP, 1 Mean, 2 ThresholdCDF(q), 3 SingleSample(seed).
 * off_m/off_v[n_off]: noise offsets sorted by m (the reference's Dict[int,int]).                              */
long long cb_scan_scratch_elems(long long n);
int cb_exclusive_scan_i32(const int* in, long long n, long long* out, long long* scratch, long long* total_out,
                          void* stream);
int cb_segment_heads(const long long* m, const int* c, long long n, int* head, int* unsorted_flag, void* stream);
int cb_segment_starts(const int* head, const long long* pos, long long n, long long* seg_start, int* seg_of,
                      void* stream);
int cb_segment_estimate(int mode, const long long* m, const int* c, const float* lp, const long long* seg_start,
                        long long n_seg, long long n_entries, const long long* off_m, const int* off_v, long long n_off,
                        double q, int n_cols, unsigned long long seed, long long* seg_m, int* out_i, float* out_f,
                        long long* map_pos, void* stream);
int cb_csr_from_segments(const long long* seg_m, long long n_seg, long long n_rows, long long n_cols, long long* indptr,
                         int* col, void* stream);
/* Output assembly after the estimators: row-wise combination of two canonical CSR matrices of equal shape,
 *     out[n,g] = row_a[n] * col_a[g] * A[n,g] + sign * row_b[n] * col_b[g] * B[n,g],  explicit zeros dropped,
 * replacing csr_set_rows_to_zero (sparse_utils.py:58-73), overwrite_matrix_with_columns_from_another
 * (sparse_utils.py:76-100), `cell_counts - noise_csr` (posterior.py:321-326) and
 * restore_eliminated_features_in_cells (data/dataset.py:495-528).  row_a[n_rows], row_b[n_rows], col_a[n_cols], col_b[n_cols]
 * are bytes (NULL = all ones); sign = +1 / -1; dtype: 0 int32, 1 int64, 2 float32, 3 float64 (both value arrays).
 * Pass 1 writes flag_a[nnz_a + 1], flag_b[nnz_b + 1] (trailing 0) and the combined values comb_a[nnz_a]; the caller
 * scans both flag arrays over nnz + 1 elements (cb_exclusive_scan_i32) and pass 2 writes out_ptr[n_rows + 1],
 * out_idx / out_val[scan_a[nnz_a] + scan_b[nnz_b]] with sorted column indices.                                      */
int cb_csr_combine_flags(long long n_rows, const long long* ptr_a, const int* idx_a, const void* val_a, long long nnz_a,
                         const long long* ptr_b, const int* idx_b, const void* val_b, long long nnz_b,
                         const unsigned char* row_a, const unsigned char* row_b, const unsigned char* col_a,
                         const unsigned char* col_b, int sign, int dtype, int* flag_a, int* flag_b, void* comb_a,
                         void* stream);
int cb_csr_combine_fill(long long n_rows, const long long* ptr_a, const int* idx_a, const long long* ptr_b,
                        const int* idx_b, const void* val_b, int sign, int dtype, const int* flag_a, const int* flag_b,
                        const long long* scan_a, const long long* scan_b, const void* comb_a, long long* out_ptr,
                        int* out_idx, void* out_val, void* stream);
/* CSR -> CSC (`.tocsc()` at posterior.py:328, dataset.py:528).  order[nnz] = permutation that stably sorts the stored
 * entries by column, sorted_col[nnz] = idx[order]; outputs col_ptr[n_cols + 1], csc_row[nnz], csc_val[nnz].          */
int cb_csr_to_csc(long long n_rows, long long n_cols, long long nnz, const long long* ptr, const void* val, int dtype,
                  const long long* order, const int* sorted_col, long long* col_ptr, int* csc_row, void* csc_val,
                  void* stream);
/* dense_to_sparse_op_torch (sparse_utils.py:10-33): nonzero flags of mask[n_rows, n_cols] (leading dimension ld_mask),
 * then, with pos = exclusive scan of flag, the (row, col, value) triplets of t in torch.nonzero's row-major order.   */
int cb_dense_nonzero_flags(const void* mask, int dtype, long long n_rows, long long n_cols, long long ld_mask, int* flag,
                           void* stream);
int cb_dense_to_coo(const void* t, int dtype, long long n_rows, long long n_cols, long long ld_t, const int* flag,
                    const long long* pos, long long* row, long long* col, void* val, void* stream);
/* PRq posterior regularisation (posterior.py:1002-1225: PRq._compute_log_target_dict, torch_binary_search over
 * PRq._get_alpha_log_constraint_violation_given_beta, renormalisation) over the (m, c)-sorted posterior COO, one
 * thread per m-segment.  log_target[n_seg] (may be NULL) = log(mean + alpha std); beta_out[n_seg] (may be NULL);
 * lp_out[n_entries] = regularised log prob (fp64 like the reference); keep[n_entries] = exp(lp_out) != 0.        */
int cb_prq_regularize(const long long* m, const int* c, const float* lp, const long long* seg_start, long long n_seg,
                      long long n_entries, const long long* off_m, const int* off_v, long long n_off, double alpha,
                      double target_tolerance, int max_iterations, double* log_target, float* beta_out, double* lp_out,
                      int* keep, void* stream);
/* PRmu's tilt with given factor(s) (posterior.py:1305-1393): beta[n_beta], n_beta = 1 (overall) or total_n_genes
 * (per-gene: beta[m % total_n_genes]); outputs as cb_prq_regularize.                                              */
int cb_posterior_tilt(const long long* m, const int* c, const float* lp, const long long* seg_start, long long n_seg,
                      long long n_entries, const long long* off_m, const int* off_v, long long n_off, const float* beta,
                      long long n_beta, long long total_n_genes, double* lp_out, int* keep, void* stream);
int cb_mckp_gene_deficit(const long long* seg_m, const int* map_val, long long n_seg, const double* target,
                         long long n_genes, long long* gene_sum, long long* deficit, void* stream);
int cb_mckp_candidates(const long long* m, const int* c, const float* lp, const int* seg_of, const long long* map_pos,
                       const long long* deficit, long long n_entries, long long n_genes, int* is_cand,
                       unsigned long long* key, void* stream);
int cb_mckp_compact(const int* is_cand, const long long* pos, const unsigned long long* key, long long n_entries,
                    unsigned long long* cand_key, long long* cand_idx, void* stream);
int cb_mckp_select(const unsigned long long* sorted_key, const long long* sorted_idx, long long n_cand,
                   const long long* deficit, const int* seg_of, int* steps, void* stream);
int cb_mckp_finalize(const long long* seg_m, const int* map_val, const int* steps, const long long* deficit,
                     long long n_seg, long long n_genes, int* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CELLBENDER_B200_H_ */
