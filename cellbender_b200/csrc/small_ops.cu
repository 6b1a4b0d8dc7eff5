// Small per-droplet pieces of the encoder and of the observation term that torch would run as dozens of 2-3 us
// elementwise launches per step (a [255]-row minibatch is pure launch latency): each is ONE launch here.
//
//   encoder_features   the hand-made features EncodeNonZLatents.forward concatenates to its inputs
//                      (reference vae/encoder.py:250-287): log1p(counts), log1p(#nonzero genes), the Gaussian overlap
//                      with the empty-droplet size, the ambient-profile dot product, ||z||_2
//   head_forward/backward   Linear(F + N -> 1) on cat(feat, x): EncodeNonZLatents.layer3 (vae/encoder.py:216, 291-293)
//                      and the last layer of eps_network (vae/encoder.py:227)
//   obs_loss / obs_small_grads   sum_b w . L and the [B]-sized gradients handed from the fused observation kernel to
//                      the latent kernel (the Dice / enumeration algebra of TraceEnum_ELBO for model.py:338-398)
#include "common.cuh"

namespace cb {

// one warp per droplet row
__global__ void __launch_bounds__(256) encoder_features_kernel(const float* __restrict__ feats, const float* __restrict__ z,
                                                               long long ldz, int B, int Z,
                                                               const float* __restrict__ d_empty_loc, float* __restrict__ p_feat,
                                                               float* __restrict__ e_feat) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  float ss = 0.f;
  for (int j = lane; j < Z; j += 32) {
    const float v = z[r * ldz + j];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float log_sum = log1pf(feats[4 * r + 0]), log_nnz = log1pf(feats[4 * r + 1]), znorm = sqrtf(ss);
    float overlap = 0.f, dot = 0.f;
    if (d_empty_loc != nullptr) {  // chi_ambient given (training / posterior); otherwise both features are zero
      const float d = fmaxf(log_sum - d_empty_loc[0], 0.f) / 0.1f;
      overlap = -0.5f * d * d;
      dot = feats[4 * r + 2];
    }
    p_feat[4 * r + 0] = log_sum, p_feat[4 * r + 1] = log_nnz, p_feat[4 * r + 2] = overlap, p_feat[4 * r + 3] = znorm;
    e_feat[4 * r + 0] = log_sum, e_feat[4 * r + 1] = log_nnz, e_feat[4 * r + 2] = dot, e_feat[4 * r + 3] = znorm;
  }
}

// out[r] = bias + sum_{j<F} w[j] feat[r, j] + sum_{j<N} w[F + j] x[r, j];  one warp per row
__global__ void __launch_bounds__(256) head_forward_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ feat,
                                                           int F, int B, int N, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  float acc = 0.f;
  for (int j = lane; j < N; j += 32) acc = fmaf(x[r * ldx + j], w[F + j], acc);
  if (feat != nullptr && lane < F) acc = fmaf(feat[r * F + lane], w[lane], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc + (bias ? bias[0] : 0.f);
}

// A CTA owns 32 of the F + N input columns (8 row groups x 32 columns, rows in registers):
//   dw[c] = sum_r d_out[r] * in[r, c];   dx[r, c - F] = d_out[r] * w[c]  (c >= F);   CTA 0 also: dbias = sum_r d_out[r]
constexpr int kHeadGroups = 8, kHeadUnroll = 32;
__global__ void __launch_bounds__(32 * kHeadGroups) head_backward_kernel(const float* __restrict__ x, long long ldx,
                                                                         const float* __restrict__ feat, int F, int B, int N,
                                                                         const float* __restrict__ w, const float* __restrict__ d_out,
                                                                         float* __restrict__ dx, long long lddx,
                                                                         float* __restrict__ dw, float* __restrict__ dbias) {
  __shared__ float s[kHeadGroups][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + c;  // column of cat(feat, x)
  const bool ok = col < F + N;
  const float wc = ok ? w[col] : 0.f;
  float acc = 0.f, accb = 0.f;
  for (int r0 = g; r0 < B; r0 += kHeadGroups * kHeadUnroll) {
    float v[kHeadUnroll], d[kHeadUnroll];
#pragma unroll
    for (int i = 0; i < kHeadUnroll; ++i) {
      const int r = r0 + i * kHeadGroups;
      const bool in = ok && r < B;
      d[i] = (r < B) ? d_out[r] : 0.f;
      v[i] = !in ? 0.f : (col < F ? feat[r * F + col] : x[r * ldx + (col - F)]);
    }
#pragma unroll
    for (int i = 0; i < kHeadUnroll; ++i) {
      const int r = r0 + i * kHeadGroups;
      acc = fmaf(d[i], v[i], acc);
      accb += d[i];
      if (dx != nullptr && ok && col >= F && r < B) dx[r * lddx + (col - F)] = d[i] * wc;
    }
  }
  s[g][c] = acc;
  __syncthreads();
  if (g == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadGroups; ++i) t += s[i][c];
    dw[col] = t;
  }
  if (dbias != nullptr && blockIdx.x == 0) {  // every column's threads saw all rows of their group: use column 0
    __syncthreads();
    s[g][c] = accb;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < kHeadGroups; ++i) t += s[i][0];
      dbias[0] = t;
    }
  }
}


// Weight gradient of the few hand-made feature columns of Linear(cat(feat, x)) (vae/encoder.py:283-298):
//   dw[h, f] = sum_b dy[b, h] * feat[b, f]     h < H, f < F <= 8, written into the first F columns of a [H, ldw] block.
// A CTA owns 32 output rows h (coalesced dy reads), its 8 warps split the minibatch rows; fixed-order reduction.
constexpr int kFeatDwGroups = 8;
__global__ void __launch_bounds__(32 * kFeatDwGroups) feat_dw_kernel(const float* __restrict__ dy, long long ldy,
                                                                      const float* __restrict__ feat, int F, int B, int H,
                                                                      float* __restrict__ dw, long long ldw) {
  __shared__ float s[kFeatDwGroups][8][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int h = blockIdx.x * 32 + c;
  float acc[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) acc[f] = 0.f;
  if (h < H) {
#pragma unroll 4
    for (int b = g; b < B; b += kFeatDwGroups) {
      const float d = dy[b * ldy + h];
#pragma unroll
      for (int f = 0; f < 8; ++f)
        if (f < F) acc[f] = fmaf(d, feat[b * F + f], acc[f]);
    }
  }
#pragma unroll
  for (int f = 0; f < 8; ++f) s[g][f][c] = acc[f];
  __syncthreads();
  if (h < H && g < F) {  // warp g finishes feature column g
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kFeatDwGroups; ++i) t += s[i][g][c];
    dw[h * ldw + g] = t;
  }
}

// loss = sum_b sum_k w[b, k] L[b, k],  neg_loss = -loss   (L = sums[:, 0:3]); one CTA, fp64 accumulation
__global__ void __launch_bounds__(256) obs_loss_kernel(const float* __restrict__ sums, int ld_sums, const float* __restrict__ w, int B,
                                                       float* __restrict__ loss, float* __restrict__ neg_loss) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* sr = sums + (long long)b * ld_sums;
    acc[0] += double(w[3 * b]) * double(sr[0]) + double(w[3 * b + 1]) * double(sr[1]) + double(w[3 * b + 2]) * double(sr[2]);
  }
  block_sum<double, 1>(acc, red);
  if (threadIdx.x == 0) {
    loss[0] = float(acc[0]);
    if (neg_loss) neg_loss[0] = float(-acc[0]);
  }
}

// gradients of loss = g * sum_b w . L with respect to the [B]-sized operands of the observation kernel:
//   d_coef[b, 0..6] = g * (w1 S3, w1 S4, w1 S5, w0 S7, w0 S8, wE S10, wE S11);  d_alpha = g * sum_b (w1 S6 + w0 S9);
//   d_w[b, 0..2] = g * L[b, 0..2]        (S* = sums[b, *] of cb_elbo_obs_fwd_bwd)
__global__ void __launch_bounds__(256) obs_small_grads_kernel(const float* __restrict__ sums, int ld_sums, const float* __restrict__ w,
                                                              const float* __restrict__ g_up, int B, float* __restrict__ d_coef,
                                                              float* __restrict__ d_alpha, float* __restrict__ d_w) {
  __shared__ double red[32];
  const float g = g_up ? g_up[0] : 1.f;
  double acc[1] = {0.0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* sr = sums + (long long)b * ld_sums;
    const float w1 = w[3 * b], w0 = w[3 * b + 1], wE = w[3 * b + 2];
    float* dc = d_coef + 7 * b;
    dc[0] = g * w1 * sr[3], dc[1] = g * w1 * sr[4], dc[2] = g * w1 * sr[5], dc[3] = g * w0 * sr[7], dc[4] = g * w0 * sr[8];
    dc[5] = g * wE * sr[10], dc[6] = g * wE * sr[11];
    if (d_w != nullptr) d_w[3 * b] = g * sr[0], d_w[3 * b + 1] = g * sr[1], d_w[3 * b + 2] = g * sr[2];
    acc[0] += double(w1) * double(sr[6]) + double(w0) * double(sr[9]);
  }
  block_sum<double, 1>(acc, red);
  if (threadIdx.x == 0) d_alpha[0] = g * float(acc[0]);
}

}  // namespace cb

extern "C" int cb_encoder_features(const float* feats, const float* z, long long ldz, int n_rows, int z_dim,
                                   const float* d_empty_loc, float* p_feat, float* e_feat, void* stream) {
  CB_REQUIRE(n_rows > 0 && z_dim >= 0, "cb_encoder_features: bad shape [%d, %d]", n_rows, z_dim);
  cb::encoder_features_kernel<<<(n_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(feats, z, ldz, n_rows, z_dim, d_empty_loc, p_feat,
                                                                                e_feat);
  return cb::check_launch("cb_encoder_features");
}

extern "C" int cb_head_forward(const float* x, long long ldx, const float* feat, int n_feat, int n_rows, int n_cols,
                               const float* w, const float* bias, float* out, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0 && n_feat >= 0 && n_feat <= 32, "cb_head_forward: bad shape [%d, %d + %d]", n_rows, n_feat,
             n_cols);
  cb::head_forward_kernel<<<(n_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, ldx, feat, n_feat, n_rows, n_cols, w, bias, out);
  return cb::check_launch("cb_head_forward");
}

extern "C" int cb_head_backward(const float* x, long long ldx, const float* feat, int n_feat, int n_rows, int n_cols,
                                const float* w, const float* d_out, float* dx, long long lddx, float* dw, float* dbias,
                                void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0 && n_feat >= 0 && n_feat <= 32, "cb_head_backward: bad shape [%d, %d + %d]", n_rows, n_feat,
             n_cols);
  const int grid = (n_feat + n_cols + 31) / 32;
  cb::head_backward_kernel<<<grid, 32 * cb::kHeadGroups, 0, (cudaStream_t)stream>>>(x, ldx, feat, n_feat, n_rows, n_cols, w, d_out, dx,
                                                                                   lddx, dw, dbias);
  return cb::check_launch("cb_head_backward");
}

extern "C" int cb_obs_loss(const float* sums, int ld_sums, const float* w, int n_rows, float* loss, float* neg_loss,
                           void* stream) {
  CB_REQUIRE(n_rows > 0 && ld_sums >= 3, "cb_obs_loss: bad shape");
  cb::obs_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, ld_sums, w, n_rows, loss, neg_loss);
  return cb::check_launch("cb_obs_loss");
}

extern "C" int cb_obs_small_grads(const float* sums, int ld_sums, const float* w, const float* g_up, int n_rows, float* d_coef,
                                  float* d_alpha, float* d_w, void* stream) {
  CB_REQUIRE(n_rows > 0 && ld_sums >= 12, "cb_obs_small_grads: bad shape");
  cb::obs_small_grads_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, ld_sums, w, g_up, n_rows, d_coef, d_alpha, d_w);
  return cb::check_launch("cb_obs_small_grads");
}

extern "C" int cb_feat_dw(const float* dy, long long ldy, const float* feat, int n_feat, int n_rows, int n_out, float* dw,
                          long long ldw, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_out > 0 && n_feat > 0 && n_feat <= 8, "cb_feat_dw: bad shape [%d, %d] x %d features", n_rows, n_out,
             n_feat);
  cb::feat_dw_kernel<<<(n_out + 31) / 32, 32 * cb::kFeatDwGroups, 0, (cudaStream_t)stream>>>(dy, ldy, feat, n_feat, n_rows, n_out, dw,
                                                                                           ldw);
  return cb::check_launch("cb_feat_dw");
}
