// BatchNorm over genes + row dropout of the EncodeNonZLatents input, forward and backward.
//
// Replaces  self.dropout50(self.batchnorm0(x))  (reference vae/encoder.py:239, 279): nn.BatchNorm1d(n_genes) on the
// dense [B, G] log-normalised counts followed by nn.Dropout1d(0.5), which on a 2-D input drops whole droplet rows
// (SURVEY.md section 8a).  torch runs this as batch_norm stats + elementwise + a 0.55 ms batch_norm_backward_reduce
// over [255, 33 538]; here one thread owns one gene column, the B rows are walked with coalesced accesses across
// the warp (the matrix is L2-resident: 34 MB << 126 MB), and nothing but the two G-sized gradient vectors is reduced.
//   training: mean/var over the minibatch (biased var for normalisation, unbiased for the running estimate),
//             y = ((x - mean) * rstd * gamma + beta) * keep[b];  running stats updated with `momentum`
//   eval:     y = (x - running_mean) * rsqrt(running_var + eps) * gamma + beta
//   backward: dgamma[g] = sum_b dy * keep[b] * xhat,  dbeta[g] = sum_b dy * keep[b]   (x is data: no input gradient)
#include "common.cuh"

namespace cb {

__global__ void __launch_bounds__(128) bn0_forward_kernel(const float* __restrict__ x, long long ldx, int B, int G,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                                         const float* __restrict__ keep,  // [B] or nullptr
                                                         int training, float momentum, float eps, float* __restrict__ y,
                                                         long long ldy, float* __restrict__ save_mean,
                                                         float* __restrict__ save_rstd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  float mean, rstd;
  if (training) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int b = 0;
    for (; b + 4 <= B; b += 4) {
      s0 += x[(long long)(b + 0) * ldx + g];
      s1 += x[(long long)(b + 1) * ldx + g];
      s2 += x[(long long)(b + 2) * ldx + g];
      s3 += x[(long long)(b + 3) * ldx + g];
    }
    for (; b < B; ++b) s0 += x[(long long)b * ldx + g];
    mean = ((s0 + s1) + (s2 + s3)) / float(B);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    b = 0;
    for (; b + 4 <= B; b += 4) {
      const float d0 = x[(long long)(b + 0) * ldx + g] - mean, d1 = x[(long long)(b + 1) * ldx + g] - mean;
      const float d2 = x[(long long)(b + 2) * ldx + g] - mean, d3 = x[(long long)(b + 3) * ldx + g] - mean;
      q0 += d0 * d0, q1 += d1 * d1, q2 += d2 * d2, q3 += d3 * d3;
    }
    for (; b < B; ++b) {
      const float d = x[(long long)b * ldx + g] - mean;
      q0 += d * d;
    }
    const float ssq = (q0 + q1) + (q2 + q3);
    const float var = ssq / float(B);
    rstd = rsqrtf(var + eps);
    running_mean[g] = (1.f - momentum) * running_mean[g] + momentum * mean;
    running_var[g] = (1.f - momentum) * running_var[g] + momentum * (B > 1 ? ssq / float(B - 1) : var);
  } else {
    mean = running_mean[g];
    rstd = rsqrtf(running_var[g] + eps);
  }
  save_mean[g] = mean;
  save_rstd[g] = rstd;
  const float a = rstd * gamma[g], c = beta[g] - mean * rstd * gamma[g];
  for (int b = 0; b < B; ++b) {
    const float k = (training && keep != nullptr) ? keep[b] : 1.f;
    y[(long long)b * ldy + g] = fmaf(x[(long long)b * ldx + g], a, c) * k;
  }
}

__global__ void __launch_bounds__(128) bn0_backward_kernel(const float* __restrict__ x, long long ldx,
                                                          const float* __restrict__ dy, long long lddy, int B, int G,
                                                          const float* __restrict__ keep, const float* __restrict__ save_mean,
                                                          const float* __restrict__ save_rstd, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const float mean = save_mean[g], rstd = save_rstd[g];
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int b = 0;
  for (; b + 2 <= B; b += 2) {
    const float k0 = keep != nullptr ? keep[b] : 1.f, k1 = keep != nullptr ? keep[b + 1] : 1.f;
    const float d0 = dy[(long long)b * lddy + g] * k0, d1 = dy[(long long)(b + 1) * lddy + g] * k1;
    a0 += d0 * (x[(long long)b * ldx + g] - mean) * rstd;
    a1 += d1 * (x[(long long)(b + 1) * ldx + g] - mean) * rstd;
    b0 += d0, b1 += d1;
  }
  for (; b < B; ++b) {
    const float d = dy[(long long)b * lddy + g] * (keep != nullptr ? keep[b] : 1.f);
    a0 += d * (x[(long long)b * ldx + g] - mean) * rstd;
    b0 += d;
  }
  dgamma[g] = a0 + a1;
  dbeta[g] = b0 + b1;
}

}  // namespace cb

extern "C" int cb_bn0_forward(const float* x, long long ldx, int n_rows, int n_genes, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, const float* keep, int training, float momentum,
                              float eps, float* y, long long ldy, float* save_mean, float* save_rstd, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_forward: bad shape [%d, %d]", n_rows, n_genes);
  cb::bn0_forward_kernel<<<(n_genes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ldx, n_rows, n_genes, gamma, beta, running_mean,
                                                                                 running_var, keep, training, momentum, eps, y, ldy,
                                                                                 save_mean, save_rstd);
  return cb::check_launch("cb_bn0_forward");
}

extern "C" int cb_bn0_backward(const float* x, long long ldx, const float* dy, long long lddy, int n_rows, int n_genes,
                               const float* keep, const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta,
                               void* stream) {
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_bn0_backward: bad shape [%d, %d]", n_rows, n_genes);
  cb::bn0_backward_kernel<<<(n_genes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ldx, dy, lddy, n_rows, n_genes, keep, save_mean,
                                                                                  save_rstd, dgamma, dbeta);
  return cb::check_launch("cb_bn0_backward");
}
