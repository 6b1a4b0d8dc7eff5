// BatchNorm1d over the minibatch (+ row dropout, + activation), forward and backward, for
//   (a) the gene-wise input normalisation of EncodeNonZLatents:  self.dropout50(self.batchnorm0(x))
//       (reference vae/encoder.py:239, 279) on the dense [B, G = 33 538] log-normalised counts, where nn.Dropout1d on a
//       2-D input drops whole droplet rows: cb_bn0_forward / cb_bn0_backward;
//   (b) the hidden layers ([B, 512]): EncodeNonZLatents.batchnorm1/2 + softplus (vae/encoder.py:283-289), eps_network
//       BatchNorm1d + Softplus (vae/encoder.py:219-228), the Decoder's FullyConnectedLayer BatchNorm1d(momentum 0.01,
//       eps 1e-3) + ReLU (vae/base.py:66-79): cb_bn_act_forward / cb_bn_act_backward.
// torch runs each of these as 4-5 launches per direction (statistics, update, transform, activation; a 0.55 ms
// batch_norm_backward_reduce over [255, 33 538]); here it is one launch each way.
//
// Layout: x, y, dy, dx are row-major [n_rows, ld].  A CTA owns 32 adjacent columns; its 256 threads are 8 row groups
// x 32 columns, so every global access is a coalesced 128-byte row segment, each thread keeps its 32 rows in registers
// (32 independent loads in flight: the kernels are latency-bound at 255 rows), and the column reductions finish in
// shared memory in a fixed order (deterministic).
//   training: mean / biased variance over the rows (two-pass), running stats updated with `momentum` (unbiased
//             variance), *num_batches_tracked += 1;  y = act(bn(x)) * keep[row]
//   eval:     running statistics, no dropout; rows are additionally split across blockIdx.y
//   act: 0 identity, 1 softplus (beta 1, threshold 20), 2 relu
//   backward: d_pre = dy * keep * act'(bn);  dbeta = sum d_pre;  dgamma = sum d_pre * xhat;
//             dx = gamma * rstd * (d_pre - dbeta / B - xhat * dgamma / B)   (training)   |   gamma * rstd * d_pre (eval)
//             dbias (optional) = column sums of dx: the gradient of a Linear bias feeding this BatchNorm
#include "common.cuh"

namespace cb {

constexpr int kBnCols = 32, kBnGroups = 8, kBnUnroll = 32;
constexpr int kBnChunk = kBnGroups * kBnUnroll;  // rows a CTA holds in registers at once (>= the 255-row minibatch)

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return v > 20.f ? v : log1pf(expf(v));
  if (act == 2) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == 1) return v > 20.f ? 1.f : 1.f / (1.f + expf(-v));
  if (act == 2) return v > 0.f ? 1.f : 0.f;
  return 1.f;
}

// sum of one value per thread over the row groups; result valid in every thread of the column
__device__ __forceinline__ float group_sum(float v, float (*s)[kBnCols + 1]) {
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  __syncthreads();
  s[g][c] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kBnGroups; ++i) t += s[i][c];
  return t;
}

// Thread (g, c) owns rows g, g + 8, g + 16, ... of column c.  When the minibatch fits one chunk (B <= 256: always on
// the training path) its rows are loaded ONCE into registers -- 32 independent loads in flight -- and statistics,
// normalisation and activation all run from there; larger B re-reads the (L2-resident) input per pass.
__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_forward_kernel(
    const float* __restrict__ x, long long ldx, int B, int N, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
    const float* __restrict__ keep, int training, float momentum, float eps, int act, float* __restrict__ y, long long ldy,
    float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float s[kBnGroups][kBnCols + 1];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnCols + c;
  const bool ok = col < N;
  // eval: this CTA normalises rows [row0, row1); training: all rows (gridDim.y == 1)
  const int rows_per_y = (B + gridDim.y - 1) / gridDim.y;
  const int row0 = blockIdx.y * rows_per_y, row1 = min(B, row0 + rows_per_y);
  const bool single = (row1 - row0) <= kBnChunk;
  float v[kBnUnroll];
  if (single) {
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      const int r = row0 + g + i * kBnGroups;
      v[i] = (ok && r < row1) ? x[(long long)r * ldx + col] : 0.f;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (training) {
    float sum = 0.f;
    if (single) {
#pragma unroll
      for (int i = 0; i < kBnUnroll; ++i) sum += v[i];  // rows beyond B hold 0
    } else if (ok) {
      for (int r = g; r < B; r += kBnGroups) sum += x[(long long)r * ldx + col];
    }
    mean = group_sum(sum, s) / float(B);
    float ssq = 0.f;
    if (single) {
#pragma unroll
      for (int i = 0; i < kBnUnroll; ++i) {
        const float d = (g + i * kBnGroups < B) ? v[i] - mean : 0.f;
        ssq += d * d;
      }
    } else if (ok) {
      for (int r = g; r < B; r += kBnGroups) {
        const float d = x[(long long)r * ldx + col] - mean;
        ssq += d * d;
      }
    }
    ssq = group_sum(ssq, s);
    const float var = ssq / float(B);
    rstd = rsqrtf(var + eps);
    if (ok && g == 0) {
      running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
      running_var[col] = (1.f - momentum) * running_var[col] + momentum * (B > 1 ? ssq / float(B - 1) : var);
    }
    if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) num_batches_tracked[0] += 1;
  } else if (ok) {
    mean = running_mean[col];
    rstd = rsqrtf(running_var[col] + eps);
  }
  if (!ok) return;
  if (g == 0 && blockIdx.y == 0) save_mean[col] = mean, save_rstd[col] = rstd;
  const float a = rstd * gamma[col], b0 = beta[col] - mean * a;
  const bool drop = training && keep != nullptr;
  if (single) {
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      const int r = row0 + g + i * kBnGroups;
      if (r < row1) y[(long long)r * ldy + col] = act_fwd(fmaf(v[i], a, b0), act) * (drop ? keep[r] : 1.f);
    }
  } else {
    for (int r = row0 + g; r < row1; r += kBnGroups)
      y[(long long)r * ldy + col] = act_fwd(fmaf(x[(long long)r * ldx + col], a, b0), act) * (drop ? keep[r] : 1.f);
  }
}

__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_backward_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int B, int N,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, const float* __restrict__ keep, int training, int act, float* __restrict__ dx,
    long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ float s[kBnGroups][kBnCols + 1];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnCols + c;
  const bool ok = col < N;
  float mean = 0.f, rstd = 0.f, gm = 0.f, bt = 0.f;
  if (ok) mean = save_mean[col], rstd = save_rstd[col], gm = gamma[col], bt = beta[col];
  const bool drop = training && keep != nullptr;
  const bool single = B <= kBnChunk;
  // xh = xhat, dp = d_pre = dy * keep * act'(bn)
  float xh[kBnUnroll], dp[kBnUnroll];
  float sb = 0.f, sg = 0.f;
  if (single) {
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      const int r = g + i * kBnGroups;
      const bool in = ok && r < B;
      xh[i] = in ? x[(long long)r * ldx + col] : mean;
      dp[i] = in ? dy[(long long)r * lddy + col] * (drop ? keep[r] : 1.f) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      xh[i] = (xh[i] - mean) * rstd;
      dp[i] *= act_grad(fmaf(xh[i], gm, bt), act);
      sb += dp[i];
      sg += dp[i] * xh[i];
    }
  } else if (ok) {
    for (int r = g; r < B; r += kBnGroups) {
      const float xhat = (x[(long long)r * ldx + col] - mean) * rstd;
      const float dpre = dy[(long long)r * lddy + col] * (drop ? keep[r] : 1.f) * act_grad(fmaf(xhat, gm, bt), act);
      sb += dpre;
      sg += dpre * xhat;
    }
  }
  sb = group_sum(sb, s);
  sg = group_sum(sg, s);
  if (ok && g == 0) dgamma[col] = sg, dbeta[col] = sb;
  if (dx == nullptr) return;  // the input is data (batchnorm0): only the affine parameters have gradients
  const float invB = 1.f / float(B);
  float sdx = 0.f;
  if (single) {
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      const int r = g + i * kBnGroups;
      if (ok && r < B) {
        const float v = training ? gm * rstd * (dp[i] - sb * invB - xh[i] * sg * invB) : gm * rstd * dp[i];
        dx[(long long)r * lddx + col] = v;
        sdx += v;
      }
    }
  } else if (ok) {
    for (int r = g; r < B; r += kBnGroups) {
      const float xhat = (x[(long long)r * ldx + col] - mean) * rstd;
      const float dpre = dy[(long long)r * lddy + col] * (drop ? keep[r] : 1.f) * act_grad(fmaf(xhat, gm, bt), act);
      const float v = training ? gm * rstd * (dpre - sb * invB - xhat * sg * invB) : gm * rstd * dpre;
      dx[(long long)r * lddx + col] = v;
      sdx += v;
    }
  }
  if (dbias) {
    sdx = group_sum(sdx, s);
    if (ok && g == 0) dbias[col] = sdx;
  }
}

static int launch_forward(const float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches_tracked, const float* keep,
                          int training, float momentum, float eps, int act, float* y, long long ldy, float* save_mean,
                          float* save_rstd, void* stream) {
  dim3 grid((n_cols + kBnCols - 1) / kBnCols, training ? 1 : (n_rows + kBnChunk - 1) / kBnChunk);
  bn_act_forward_kernel<<<grid, kBnCols * kBnGroups, 0, (cudaStream_t)stream>>>(
      x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, keep, training, momentum, eps, act, y,
      ldy, save_mean, save_rstd);
  return 0;
}

}  // namespace cb

extern "C" int cb_bn_act_forward(const float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, long long* num_batches_tracked, int training,
                                 float momentum, float eps, int act, float* y, long long ldy, float* save_mean,
                                 float* save_rstd, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_forward: empty input");
  CB_REQUIRE(act >= 0 && act <= 2, "cb_bn_act_forward: act must be 0 (identity), 1 (softplus) or 2 (relu)");
  cb::launch_forward(x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, nullptr, training,
                     momentum, eps, act, y, ldy, save_mean, save_rstd, stream);
  return cb::check_launch("cb_bn_act_forward");
}

extern "C" int cb_bn_act_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_cols,
                                  const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
                                  int training, int act, float* dx, long long lddx, float* dgamma, float* dbeta, float* dbias,
                                  void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_backward: empty input");
  const int grid = (n_cols + cb::kBnCols - 1) / cb::kBnCols;
  cb::bn_act_backward_kernel<<<grid, cb::kBnCols * cb::kBnGroups, 0, (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, n_rows, n_cols, gamma, beta, save_mean, save_rstd, nullptr, training, act, dx, lddx, dgamma, dbeta, dbias);
  return cb::check_launch("cb_bn_act_backward");
}

extern "C" int cb_bn0_forward(const float* x, long long ldx, int n_rows, int n_genes, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, const float* keep, int training, float momentum,
                              float eps, float* y, long long ldy, float* save_mean, float* save_rstd, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_forward: bad shape [%d, %d]", n_rows, n_genes);
  cb::launch_forward(x, ldx, n_rows, n_genes, gamma, beta, running_mean, running_var, nullptr, keep, training, momentum, eps, 0,
                     y, ldy, save_mean, save_rstd, stream);
  return cb::check_launch("cb_bn0_forward");
}

extern "C" int cb_bn0_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_genes,
                               const float* keep, const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta,
                               void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_backward: bad shape [%d, %d]", n_rows, n_genes);
  const int grid = (n_genes + cb::kBnCols - 1) / cb::kBnCols;
  // gamma / beta only enter through act'(.), which is 1 for the identity: pass the statistics as dummies
  cb::bn_act_backward_kernel<<<grid, cb::kBnCols * cb::kBnGroups, 0, (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, n_rows, n_genes, save_mean, save_mean, save_mean, save_rstd, keep, 1, 0, nullptr, 0, dgamma, dbeta, nullptr);
  return cb::check_launch("cb_bn0_backward");
}
