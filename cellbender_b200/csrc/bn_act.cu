// BatchNorm1d over the minibatch + activation for the hidden layers ([B, 512]-shaped), forward and backward.
//
// Replaces the  nn.BatchNorm1d -> nn.Softplus / nn.ReLU  pairs of the reference's hidden layers
// (EncodeNonZLatents.batchnorm1/2 + softplus, vae/encoder.py:283-289; eps_network BatchNorm1d + Softplus,
// vae/encoder.py:219-228; FullyConnectedLayer of the Decoder: BatchNorm1d(momentum 0.01, eps 1e-3) + ReLU,
// vae/base.py:66-79) and their autograd backward: torch runs each pair as 5 forward + 4 backward kernels, each a few
// microseconds of launch latency on a 0.5 MB tensor.  Here: one launch each way.
//
// Layout: x, y, dy, dx are row-major [n_rows, n_cols] with leading dimension ld.  A CTA owns 32 adjacent columns;
// its 256 threads are 8 row groups x 32 columns, so every global access is a coalesced 128-byte row segment and the
// column reductions finish in shared memory (fixed order: deterministic).
//   training: mean / biased variance over the rows (two-pass), running stats updated with `momentum`
//             (unbiased variance), *num_batches_tracked += 1
//   eval:     running statistics
//   act: 0 identity, 1 softplus (beta 1, threshold 20), 2 relu
//   backward: d_pre = dy * act'(bn);  dbeta = sum d_pre;  dgamma = sum d_pre * xhat;
//             dx = gamma * rstd * (d_pre - dbeta / B - xhat * dgamma / B)   (training)   |   gamma * rstd * d_pre (eval)
//             dbias (optional) = column sums of dx: the gradient of a Linear bias feeding this BatchNorm
#include "common.cuh"

namespace cb {

constexpr int kBnCols = 32, kBnGroups = 8;

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

// sum over the 8 row groups of one value per thread; result valid for every thread of the column
__device__ __forceinline__ float group_sum(float v, float (*s)[kBnCols]) {
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  __syncthreads();
  s[g][c] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kBnGroups; ++i) t += s[i][c];
  return t;
}

__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_forward_kernel(
    const float* __restrict__ x, long long ldx, int B, int N, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
    int training, float momentum, float eps, int act, float* __restrict__ y, long long ldy, float* __restrict__ save_mean,
    float* __restrict__ save_rstd) {
  __shared__ float s[kBnGroups][kBnCols];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnCols + c;
  const bool ok = col < N;
  float mean = 0.f, rstd = 1.f;
  if (training) {
    float sum = 0.f;
    if (ok)
      for (int r = g; r < B; r += kBnGroups) sum += x[(long long)r * ldx + col];
    mean = group_sum(sum, s) / float(B);
    float ssq = 0.f;
    if (ok)
      for (int r = g; r < B; r += kBnGroups) {
        const float d = x[(long long)r * ldx + col] - mean;
        ssq += d * d;
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
  if (g == 0) save_mean[col] = mean, save_rstd[col] = rstd;
  const float a = rstd * gamma[col], b0 = beta[col] - mean * a;
  for (int r = g; r < B; r += kBnGroups) y[(long long)r * ldy + col] = act_fwd(fmaf(x[(long long)r * ldx + col], a, b0), act);
}

__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_backward_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int B, int N,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, int training, int act, float* __restrict__ dx, long long lddx,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ float s[kBnGroups][kBnCols];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int col = blockIdx.x * kBnCols + c;
  const bool ok = col < N;
  float mean = 0.f, rstd = 0.f, gm = 0.f, bt = 0.f;
  if (ok) mean = save_mean[col], rstd = save_rstd[col], gm = gamma[col], bt = beta[col];
  float sb = 0.f, sg = 0.f;
  if (ok)
    for (int r = g; r < B; r += kBnGroups) {
      const float xhat = (x[(long long)r * ldx + col] - mean) * rstd;
      const float dpre = dy[(long long)r * lddy + col] * act_grad(fmaf(xhat, gm, bt), act);
      sb += dpre;
      sg += dpre * xhat;
    }
  sb = group_sum(sb, s);
  sg = group_sum(sg, s);
  const float invB = 1.f / float(B);
  float sdx = 0.f;
  if (ok)
    for (int r = g; r < B; r += kBnGroups) {
      const float xhat = (x[(long long)r * ldx + col] - mean) * rstd;
      const float dpre = dy[(long long)r * lddy + col] * act_grad(fmaf(xhat, gm, bt), act);
      const float v = training ? gm * rstd * (dpre - sb * invB - xhat * sg * invB) : gm * rstd * dpre;
      dx[(long long)r * lddx + col] = v;
      sdx += v;
    }
  if (dbias) sdx = group_sum(sdx, s);
  if (ok && g == 0) {
    dgamma[col] = sg;
    dbeta[col] = sb;
    if (dbias) dbias[col] = sdx;
  }
}

}  // namespace cb

extern "C" int cb_bn_act_forward(const float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, long long* num_batches_tracked, int training,
                                 float momentum, float eps, int act, float* y, long long ldy, float* save_mean,
                                 float* save_rstd, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_forward: empty input");
  CB_REQUIRE(act >= 0 && act <= 2, "cb_bn_act_forward: act must be 0 (identity), 1 (softplus) or 2 (relu)");
  const int grid = (n_cols + cb::kBnCols - 1) / cb::kBnCols;
  cb::bn_act_forward_kernel<<<grid, cb::kBnCols * cb::kBnGroups, 0, (cudaStream_t)stream>>>(
      x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, training, momentum, eps, act, y, ldy,
      save_mean, save_rstd);
  return cb::check_launch("cb_bn_act_forward");
}

extern "C" int cb_bn_act_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_cols,
                                  const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
                                  int training, int act, float* dx, long long lddx, float* dgamma, float* dbeta, float* dbias,
                                  void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_backward: empty input");
  const int grid = (n_cols + cb::kBnCols - 1) / cb::kBnCols;
  cb::bn_act_backward_kernel<<<grid, cb::kBnCols * cb::kBnGroups, 0, (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, n_rows, n_cols, gamma, beta, save_mean, save_rstd, training, act, dx, lddx, dgamma, dbeta, dbias);
  return cb::check_launch("cb_bn_act_backward");
}
