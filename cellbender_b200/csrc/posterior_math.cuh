// fp32 device math for the posterior noise-count pmf of ONE (cell, gene) entry and ONE latent sample.
//
// Replaces Posterior._log_prob_noise_count_tensor + the logsumexp normalisation of noise_log_pdf
// (reference: cellbender/remove_background/posterior.py:815-859 and :716-718), evaluated only where
// the count x > 0 (the reference discards x == 0 entries afterwards, posterior.py:532).
//
// Reference per candidate noise count c in [low, low + n):
//   lp(c) = Poisson(c; lam) + NegBinomial(x - c; total_count = alpha, logits = log mu - log alpha),  -inf if c > x
// and then lp(c) - logsumexp_c lp(c).  Every factor that does not depend on c cancels in the
// normalisation, and consecutive terms differ by a rational factor (no lgamma needed):
//   p(c+1)/p(c) = lam (x - c) / ((c + 1) (alpha + x - c - 1) s),   s = sigmoid(logits) = mu / (mu + alpha)
// so the normalised window is computed from n - 1 logs, exactly in the sense of exact arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef CB_HD
#define CB_HD __host__ __device__ __forceinline__
#endif

namespace cb {

// Window start (posterior.py:816-819): low = clamp(min(int(lam - n/2), int(x - n + 1)), 0); int() truncates.
CB_HD int noise_window_low(float x, float lam, int n) {
  const int a = int(lam - 0.5f * float(n));
  const int b = int(x - float(n) + 1.0f);
  const int lo = a < b ? a : b;
  return lo > 0 ? lo : 0;
}

// Normalised pmf over the window.  prob[j] = P(noise = low + j), j < n  (0 where low + j > x).
// Returns low.  mu, lam, alpha must already carry the reference's +1e-30 safeguards (posterior.py:708-710).
template <int NMAX>
CB_HD int noise_window_pmf(float x, float mu, float lam, float alpha, int n, float (&prob)[NMAX]) {
  const int low = noise_window_low(x, lam, n);
  // log s = logsigmoid(log mu - log alpha) = -log1p(alpha / mu); alpha/mu may overflow to inf -> -inf handled below
  const float log_lam_over_s = logf(lam) + log1pf(alpha / mu);
  float t[NMAX];
  float tmax = 0.0f;
  t[0] = 0.0f;  // relative log prob of c = low (always <= x because low <= max(x - n + 1, 0) <= x)
  bool dead = false;
#pragma unroll
  for (int j = 1; j < NMAX; ++j) {
    if (j < n) {
      const float c = float(low + j - 1);  // ratio p(c + 1) / p(c)
      const float k = x - c;               // NB count at c
      if (dead || k <= 0.0f) {
        dead = true;
        t[j] = -INFINITY;
      } else {
        t[j] = t[j - 1] + (log_lam_over_s + logf(k / ((c + 1.0f) * (alpha + k - 1.0f))));
        tmax = fmaxf(tmax, t[j]);
      }
    } else {
      t[j] = -INFINITY;
    }
  }
  // mu -> 0 (+1e-30): log_lam_over_s = +inf -> mass piles on the last valid c; handle inf - inf
  if (isinf(tmax)) {
    int last = 0;
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < n && float(low + j) <= x) last = j;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) prob[j] = (j == last) ? 1.0f : 0.0f;
    return low;
  }
  float z = 0.0f;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    prob[j] = (j < n) ? expf(t[j] - tmax) : 0.0f;
    z += prob[j];
  }
  const float inv = 1.0f / z;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) prob[j] *= inv;
  return low;
}

// The same pmf accumulated in place: acc[j] += P(noise = low + j).  Arithmetic identical to noise_window_pmf (same
// operations in the same order); only the temporaries are shared (one NMAX-array instead of two) so that the sample
// loop of the posterior kernel keeps acc[] + t[] = 2 NMAX registers live.  Returns low.
template <int NMAX>
CB_HD int noise_window_pmf_accumulate(float x, float mu, float lam, float alpha, int n, float (&acc)[NMAX]) {
  const int low = noise_window_low(x, lam, n);
  const float log_lam_over_s = logf(lam) + log1pf(alpha / mu);
  float t[NMAX];
  float tmax = 0.0f;
  t[0] = 0.0f;
  bool dead = false;
#pragma unroll
  for (int j = 1; j < NMAX; ++j) {
    if (j < n) {
      const float c = float(low + j - 1);
      const float k = x - c;
      if (dead || k <= 0.0f) {
        dead = true;
        t[j] = -INFINITY;
      } else {
        t[j] = t[j - 1] + (log_lam_over_s + logf(k / ((c + 1.0f) * (alpha + k - 1.0f))));
        tmax = fmaxf(tmax, t[j]);
      }
    } else {
      t[j] = -INFINITY;
    }
  }
  if (isinf(tmax)) {
    int last = 0;
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < n && float(low + j) <= x) last = j;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] += (j == last) ? 1.0f : 0.0f;
    return low;
  }
  float z = 0.0f;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    t[j] = (j < n) ? expf(t[j] - tmax) : 0.0f;
    z += t[j];
  }
  const float inv = 1.0f / z;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) acc[j] += t[j] * inv;
  return low;
}

}  // namespace cb
