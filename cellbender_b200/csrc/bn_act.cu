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

// Two shapes of the same kernels (COLS adjacent columns x GROUPS row groups = 256 threads, UNROLL rows per thread in
// registers, GROUPS * UNROLL = 256 >= the 255-row minibatch):
//   wide   <32, 8, 32>  128-byte row segments: the gene-wide matrices (thousands of column strips fill the GPU)
//   narrow <8, 32, 8>   32-byte row segments (one sector): the [B, 512] hidden layers, where the wide shape gives only
//                       16 CTAs of 127 instructions per element each (ncu r02g: 14-29 us, issue-bound on 16 SMs);
//                       64 CTAs with a quarter of the per-thread work cut that chain link to a few microseconds.
constexpr int kBnChunk = 256;  // rows a CTA holds in registers at once

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
template <int COLS, int GROUPS>
__device__ __forceinline__ float group_sum(float v, float (*s)[COLS + 1]) {
  const int c = threadIdx.x % COLS, g = threadIdx.x / COLS;
  __syncthreads();
  s[g][c] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < GROUPS; ++i) t += s[i][c];
  return t;
}

// Thread (g, c) owns rows g, g + 8, g + 16, ... of column c.  When the minibatch fits one chunk (B <= 256: always on
// the training path) its rows are loaded ONCE into registers -- 32 independent loads in flight -- and statistics,
// normalisation and activation all run from there; larger B re-reads the (L2-resident) input per pass.
//
// Optional rank-F feature term (F <= 8): the input is x[r, c] + sum_f feat[r, f] * wf[c, f] -- the hand-made feature columns
// of Linear(cat(feat, x_in)) (vae/encoder.py:283-298).  The G-wide gene block of that Linear is produced by the tensor-core
// GEMM as soon as batchnorm0 is done; the features need z (EncodeZ) and join here, so the big products do not wait for
// them.  The completed pre-activation is written back to x (the backward kernel re-reads it).
template <int kBnCols, int kBnGroups, int kBnUnroll>
__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_forward_kernel(
    float* __restrict__ x, long long ldx, int B, int N, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
    const float* __restrict__ keep, int training, float momentum, float eps, int act, float* __restrict__ y, long long ldy,
    float* __restrict__ save_mean, float* __restrict__ save_rstd, const float* __restrict__ feat, int F,
    const float* __restrict__ wf, long long ldw) {
  static_assert(kBnGroups * kBnUnroll == kBnChunk, "a CTA pass covers kBnChunk rows");
  __shared__ float s[kBnGroups][kBnCols + 1];
  const int c = threadIdx.x % kBnCols, g = threadIdx.x / kBnCols;
  const int col = blockIdx.x * kBnCols + c;
  const bool ok = col < N;
  // eval: this CTA normalises rows [row0, row1); training: all rows (gridDim.y == 1)
  const int rows_per_y = (B + gridDim.y - 1) / gridDim.y;
  const int row0 = blockIdx.y * rows_per_y, row1 = min(B, row0 + rows_per_y);
  const bool single = (row1 - row0) <= kBnChunk;
  float wfc[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) wfc[f] = (feat != nullptr && ok && f < F) ? wf[(long long)col * ldw + f] : 0.f;
  float v[kBnUnroll];
  if (single) {
#pragma unroll
    for (int i = 0; i < kBnUnroll; ++i) {
      const int r = row0 + g + i * kBnGroups;
      v[i] = (ok && r < row1) ? x[(long long)r * ldx + col] : 0.f;
    }
    if (feat != nullptr) {
#pragma unroll
      for (int i = 0; i < kBnUnroll; ++i) {
        const int r = row0 + g + i * kBnGroups;
        if (ok && r < row1) {
#pragma unroll
          for (int f = 0; f < 8; ++f)
            if (f < F) v[i] = fmaf(feat[r * F + f], wfc[f], v[i]);
          x[(long long)r * ldx + col] = v[i];
        }
      }
    }
  } else if (feat != nullptr && ok) {
    // rows of this thread only: it re-reads exactly what it wrote (training re-reads ALL its own rows: row0 = 0, row1 = B)
    for (int r = row0 + g; r < row1; r += kBnGroups) {
      float t = x[(long long)r * ldx + col];
#pragma unroll
      for (int f = 0; f < 8; ++f)
        if (f < F) t = fmaf(feat[r * F + f], wfc[f], t);
      x[(long long)r * ldx + col] = t;
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
    mean = group_sum<kBnCols, kBnGroups>(sum, s) / float(B);
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
    ssq = group_sum<kBnCols, kBnGroups>(ssq, s);
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

template <int kBnCols, int kBnGroups, int kBnUnroll>
__global__ void __launch_bounds__(kBnCols* kBnGroups) bn_act_backward_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int B, int N,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_rstd, const float* __restrict__ keep, int training, int act, float* __restrict__ dx,
    long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ float s[kBnGroups][kBnCols + 1];
  const int c = threadIdx.x % kBnCols, g = threadIdx.x / kBnCols;
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
  sb = group_sum<kBnCols, kBnGroups>(sb, s);
  sg = group_sum<kBnCols, kBnGroups>(sg, s);
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
    sdx = group_sum<kBnCols, kBnGroups>(sdx, s);
    if (ok && g == 0) dbias[col] = sdx;
  }
}

// ---- batchnorm0 over the wide [B, G] input: float4 column strips ---------------------------------------------------------
// A CTA owns 64 adjacent columns (256 contiguous bytes per row): 16 row groups x 16 lanes x float4.  The generic kernel
// above moves 128-byte row segments with 4-byte accesses and reached 1.6 TB/s (forward) / 1.0 TB/s (backward) on the
// 69 MB input (r02 timeline: 44 / 70 us); these move 16 bytes per access and keep 16 float4 loads per thread in flight.
// Forward: the thread's <= 16 rows stay in registers for the two-pass statistics and the normalisation.  Backward: the
// input is data, only dgamma / dbeta exist -- a pure streaming column reduction, nothing is kept.
constexpr int kB0Lanes = 16, kB0Groups = 16, kB0Rows = 16;  // rows per thread held in registers (forward)
constexpr int kB0Chunk = kB0Groups * kB0Rows;                // 256 rows per CTA pass

__device__ __forceinline__ float4 group_sum4(float4 v, float4 (*s)[kB0Lanes + 1]) {
  const int c = threadIdx.x & (kB0Lanes - 1), g = threadIdx.x / kB0Lanes;
  __syncthreads();
  s[g][c] = v;
  __syncthreads();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kB0Groups; ++i) {
    const float4 u = s[i][c];
    t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
  }
  return t;
}

__global__ void __launch_bounds__(kB0Lanes* kB0Groups) bn0_forward_kernel(
    const float* __restrict__ x, long long ldx, int B, int G, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
    const float* __restrict__ keep, int training, float momentum, float eps, float* __restrict__ y, long long ldy,
    float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float4 s[kB0Groups][kB0Lanes + 1];
  if (training && num_batches_tracked != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    num_batches_tracked[0] += 1;  // nn.BatchNorm1d's counter (a separate int64 add-on-self launch sat in front of this kernel)
  const int c = threadIdx.x & (kB0Lanes - 1), g = threadIdx.x / kB0Lanes;
  const int col = (blockIdx.x * kB0Lanes + c) * 4;
  const bool ok = col < G;  // the float4 may straddle G: ld is padded to a multiple of 4, lanes beyond G are masked below
  const int rows_per_y = (B + gridDim.y - 1) / gridDim.y;
  const int row0 = blockIdx.y * rows_per_y, row1 = min(B, row0 + rows_per_y);  // training: gridDim.y == 1, all rows
  float4 v[kB0Rows];
#pragma unroll
  for (int i = 0; i < kB0Rows; ++i) {
    const int r = row0 + g + i * kB0Groups;
    v[i] = (ok && r < row1) ? *reinterpret_cast<const float4*>(x + (long long)r * ldx + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), rstd = make_float4(1.f, 1.f, 1.f, 1.f);
  const bool in[4] = {col < G, col + 1 < G, col + 2 < G, col + 3 < G};
  if (training) {
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kB0Rows; ++i) sum.x += v[i].x, sum.y += v[i].y, sum.z += v[i].z, sum.w += v[i].w;
    sum = group_sum4(sum, s);
    const float invB = 1.f / float(B);
    mean = make_float4(sum.x * invB, sum.y * invB, sum.z * invB, sum.w * invB);
    float4 ssq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kB0Rows; ++i) {
      if (g + i * kB0Groups < B) {
        const float dx_ = v[i].x - mean.x, dy_ = v[i].y - mean.y, dz_ = v[i].z - mean.z, dw_ = v[i].w - mean.w;
        ssq.x += dx_ * dx_, ssq.y += dy_ * dy_, ssq.z += dz_ * dz_, ssq.w += dw_ * dw_;
      }
    }
    ssq = group_sum4(ssq, s);
    rstd = make_float4(rsqrtf(ssq.x * invB + eps), rsqrtf(ssq.y * invB + eps), rsqrtf(ssq.z * invB + eps), rsqrtf(ssq.w * invB + eps));
    if (ok && g == 0) {
      const float m4[4] = {mean.x, mean.y, mean.z, mean.w}, q4[4] = {ssq.x, ssq.y, ssq.z, ssq.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (in[k]) {
          running_mean[col + k] = (1.f - momentum) * running_mean[col + k] + momentum * m4[k];
          running_var[col + k] = (1.f - momentum) * running_var[col + k] + momentum * (B > 1 ? q4[k] / float(B - 1) : q4[k] * invB);
        }
    }
  } else if (ok) {
    float m4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (in[k]) m4[k] = running_mean[col + k], r4[k] = rsqrtf(running_var[col + k] + eps);
    mean = make_float4(m4[0], m4[1], m4[2], m4[3]), rstd = make_float4(r4[0], r4[1], r4[2], r4[3]);
  }
  if (!ok) return;
  float a4[4], b4[4];
  {
    const float m4[4] = {mean.x, mean.y, mean.z, mean.w}, r4[4] = {rstd.x, rstd.y, rstd.z, rstd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gm = in[k] ? gamma[col + k] : 0.f, bt = in[k] ? beta[col + k] : 0.f;
      a4[k] = r4[k] * gm, b4[k] = bt - m4[k] * a4[k];
      if (g == 0 && blockIdx.y == 0 && in[k]) save_mean[col + k] = m4[k], save_rstd[col + k] = r4[k];
    }
  }
  const bool drop = training && keep != nullptr;
#pragma unroll
  for (int i = 0; i < kB0Rows; ++i) {
    const int r = row0 + g + i * kB0Groups;
    if (r < row1) {
      const float k_ = drop ? keep[r] : 1.f;
      float4 o;
      o.x = in[0] ? fmaf(v[i].x, a4[0], b4[0]) * k_ : 0.f;
      o.y = in[1] ? fmaf(v[i].y, a4[1], b4[1]) * k_ : 0.f;
      o.z = in[2] ? fmaf(v[i].z, a4[2], b4[2]) * k_ : 0.f;
      o.w = in[3] ? fmaf(v[i].w, a4[3], b4[3]) * k_ : 0.f;
      *reinterpret_cast<float4*>(y + (long long)r * ldy + col) = o;
    }
  }
}

// dgamma[c] = sum_r dy[r, c] keep[r] xhat[r, c],  dbeta[c] = sum_r dy[r, c] keep[r]   (identity activation, no dx)
__global__ void __launch_bounds__(kB0Lanes* kB0Groups) bn0_backward_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int B, int G,
    const float* __restrict__ keep, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float4 s[kB0Groups][kB0Lanes + 1];
  const int c = threadIdx.x & (kB0Lanes - 1), g = threadIdx.x / kB0Lanes;
  const int col = (blockIdx.x * kB0Lanes + c) * 4;
  const bool ok = col < G;
  const bool in[4] = {col < G, col + 1 < G, col + 2 < G, col + 3 < G};
  float m4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (in[k]) m4[k] = save_mean[col + k], r4[k] = save_rstd[col + k];
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    for (int r0 = g; r0 < B; r0 += kB0Groups * 8) {
      float4 xv[8], dv[8];
      float kk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // all loads first
        const int r = r0 + u * kB0Groups;
        const bool rin = r < B;
        xv[u] = rin ? *reinterpret_cast<const float4*>(x + (long long)r * ldx + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[u] = rin ? *reinterpret_cast<const float4*>(dy + (long long)r * lddy + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        kk[u] = rin ? (keep != nullptr ? keep[r] : 1.f) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float dx_ = dv[u].x * kk[u], dy_ = dv[u].y * kk[u], dz_ = dv[u].z * kk[u], dw_ = dv[u].w * kk[u];
        sb.x += dx_, sb.y += dy_, sb.z += dz_, sb.w += dw_;
        sg.x += dx_ * ((xv[u].x - m4[0]) * r4[0]), sg.y += dy_ * ((xv[u].y - m4[1]) * r4[1]);
        sg.z += dz_ * ((xv[u].z - m4[2]) * r4[2]), sg.w += dw_ * ((xv[u].w - m4[3]) * r4[3]);
      }
    }
  }
  sb = group_sum4(sb, s);
  sg = group_sum4(sg, s);
  if (ok && g == 0) {
    const float b4[4] = {sb.x, sb.y, sb.z, sb.w}, g4[4] = {sg.x, sg.y, sg.z, sg.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (in[k]) dgamma[col + k] = g4[k], dbeta[col + k] = b4[k];
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int launch_forward(float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches_tracked, const float* keep,
                          int training, float momentum, float eps, int act, float* y, long long ldy, float* save_mean,
                          float* save_rstd, const float* feat, int n_feat, const float* wf, long long ldw, void* stream) {
  const int gy = training ? 1 : (n_rows + kBnChunk - 1) / kBnChunk;
  if (n_cols <= 2048) {  // narrow shape: four times the CTAs, a quarter of the rows per thread
    dim3 grid((n_cols + 7) / 8, gy);
    bn_act_forward_kernel<8, 32, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(
        x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, keep, training, momentum, eps, act, y,
        ldy, save_mean, save_rstd, feat, n_feat, wf, ldw);
  } else {
    dim3 grid((n_cols + 31) / 32, gy);
    bn_act_forward_kernel<32, 8, 32><<<grid, 256, 0, (cudaStream_t)stream>>>(
        x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, keep, training, momentum, eps, act, y,
        ldy, save_mean, save_rstd, feat, n_feat, wf, ldw);
  }
  return 0;
}

}  // namespace cb

extern "C" int cb_bn_act_forward(float* x, long long ldx, int n_rows, int n_cols, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, long long* num_batches_tracked, int training,
                                 float momentum, float eps, int act, float* y, long long ldy, float* save_mean,
                                 float* save_rstd, const float* feat, int n_feat, const float* feat_weight, long long ldw,
                                 void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_forward: empty input");
  CB_REQUIRE(act >= 0 && act <= 2, "cb_bn_act_forward: act must be 0 (identity), 1 (softplus) or 2 (relu)");
  CB_REQUIRE(feat == nullptr || (n_feat > 0 && n_feat <= 8 && feat_weight != nullptr && ldw >= n_feat),
             "cb_bn_act_forward: the feature term needs 1..8 features and their weight columns");
  cb::launch_forward(x, ldx, n_rows, n_cols, gamma, beta, running_mean, running_var, num_batches_tracked, nullptr, training,
                     momentum, eps, act, y, ldy, save_mean, save_rstd, feat, feat == nullptr ? 0 : n_feat, feat_weight, ldw, stream);
  return cb::check_launch("cb_bn_act_forward");
}

extern "C" int cb_bn_act_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_cols,
                                  const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
                                  int training, int act, float* dx, long long lddx, float* dgamma, float* dbeta, float* dbias,
                                  void* stream) {
  CB_REQUIRE(n_rows > 0 && n_cols > 0, "cb_bn_act_backward: empty input");
  if (n_cols <= 2048)
    cb::bn_act_backward_kernel<8, 32, 8><<<(n_cols + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        x, ldx, dy, lddy, n_rows, n_cols, gamma, beta, save_mean, save_rstd, nullptr, training, act, dx, lddx, dgamma, dbeta, dbias);
  else
    cb::bn_act_backward_kernel<32, 8, 32><<<(n_cols + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
        x, ldx, dy, lddy, n_rows, n_cols, gamma, beta, save_mean, save_rstd, nullptr, training, act, dx, lddx, dgamma, dbeta, dbias);
  return cb::check_launch("cb_bn_act_backward");
}

extern "C" int cb_bn0_forward(const float* x, long long ldx, int n_rows, int n_genes, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, long long* num_batches_tracked, const float* keep,
                              int training, float momentum, float eps, float* y, long long ldy, float* save_mean,
                              float* save_rstd, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_forward: bad shape [%d, %d]", n_rows, n_genes);
  if ((ldx & 3) == 0 && (ldy & 3) == 0 && cb::aligned16(x) && cb::aligned16(y) && (!training || n_rows <= cb::kB0Chunk)) {
    dim3 grid((n_genes + 4 * cb::kB0Lanes - 1) / (4 * cb::kB0Lanes), training ? 1 : (n_rows + cb::kB0Chunk - 1) / cb::kB0Chunk);
    cb::bn0_forward_kernel<<<grid, cb::kB0Lanes * cb::kB0Groups, 0, (cudaStream_t)stream>>>(
        x, ldx, n_rows, n_genes, gamma, beta, running_mean, running_var, num_batches_tracked, keep, training, momentum, eps, y, ldy,
        save_mean, save_rstd);
    return cb::check_launch("cb_bn0_forward");
  }
  cb::launch_forward(const_cast<float*>(x), ldx, n_rows, n_genes, gamma, beta, running_mean, running_var,
                     training ? num_batches_tracked : nullptr, keep, training,
                     momentum, eps, 0, y, ldy, save_mean, save_rstd, nullptr, 0, nullptr, 0, stream);
  return cb::check_launch("cb_bn0_forward");
}

extern "C" int cb_bn0_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_genes,
                               const float* keep, const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta,
                               void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_backward: bad shape [%d, %d]", n_rows, n_genes);
  if ((ldx & 3) == 0 && (lddy & 3) == 0 && cb::aligned16(x) && cb::aligned16(dy)) {
    const int grid4 = (n_genes + 4 * cb::kB0Lanes - 1) / (4 * cb::kB0Lanes);
    cb::bn0_backward_kernel<<<grid4, cb::kB0Lanes * cb::kB0Groups, 0, (cudaStream_t)stream>>>(x, ldx, dy, lddy, n_rows, n_genes, keep,
                                                                                           save_mean, save_rstd, dgamma, dbeta);
    return cb::check_launch("cb_bn0_backward");
  }
  const int grid = (n_genes + 31) / 32;
  // gamma / beta only enter through act'(.), which is 1 for the identity: pass the statistics as dummies
  cb::bn_act_backward_kernel<32, 8, 32><<<grid, 256, 0, (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, n_rows, n_genes, save_mean, save_mean, save_mean, save_rstd, keep, 1, 0, nullptr, 0, dgamma, dbeta, nullptr);
  return cb::check_launch("cb_bn0_backward");
}
