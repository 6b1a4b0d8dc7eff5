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
// so the normalised window is a running product of n - 1 rational factors, exactly in the sense of exact arithmetic.
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

// Normalised pmf over the window, accumulated in place: acc[j] += P(noise = low + j), j < n (0 where low + j > x).
// mu, lam, alpha must already carry the reference's +1e-30 safeguards (posterior.py:708-710).  No transcendental functions: consecutive terms
// differ by the rational factor r(c) = (lam / s) (x - c) / ((c + 1) (alpha + x - c - 1)), lam / s = lam (1 + alpha / mu), so
// the un-normalised pmf is a running product.  Its dynamic range is tamed by rescaling with exact powers of two whenever
// the product leaves [2^-64, 2^64] (the rescale events are remembered as two bit masks), and r is clamped to
// [2^-60, 2^60] -- a neighbour more than 10^18 times less likely lies far below every threshold the posterior is stored
// with (log p >= -10).  Per sample this is n divisions and multiplications instead of ~n logf + n expf: the window
// kernel was issue-bound on exactly those (ncu r02k: 980 M warp instructions per chunk, 1 240 per sample evaluation).
// Relative error of the result ~ n x 2^-24 (products) -- tighter than the summed-logarithm form.  Returns low.
template <int NMAX>
CB_HD int noise_window_pmf_accumulate(float x, float mu, float lam, float alpha, int n, float (&acc)[NMAX]) {
  static_assert(NMAX <= 64, "rescale masks are 64-bit");
  const int low = noise_window_low(x, lam, n);
  const float a_ratio = fminf(lam * (1.0f + alpha / mu), 3.0e37f);  // lam / s; alpha / mu may overflow: clamp, r is clamped anyway
  constexpr float kBig = 18446744073709551616.0f, kSmall = 5.421010862427522e-20f;  // 2^64, 2^-64
  constexpr float kRmax = 1152921504606846976.0f, kRmin = 8.673617379884035e-19f;   // 2^60, 2^-60
  float w[NMAX];
  unsigned long long up = 0ull, down = 0ull;
  float cur = 1.0f;
  w[0] = 1.0f;  // relative probability of c = low (always <= x because low <= max(x - n + 1, 0) <= x)
  int scale = 0, best_scale = 0;  // in units of 2^64
  float best = 1.0f;
  bool dead = false;
#pragma unroll
  for (int j = 1; j < NMAX; ++j) {
    float wj = 0.0f;
    if (j < n && !dead) {
      const float c = float(low + j - 1);  // ratio p(c + 1) / p(c)
      const float k = x - c;               // NB count at c
      if (k <= 0.0f) {
        dead = true;
      } else {
        float r = a_ratio * (k / ((c + 1.0f) * (alpha + k - 1.0f)));
        r = fminf(fmaxf(r, kRmin), kRmax);
        cur *= r;
        if (cur > kBig) {
          cur *= kSmall, up |= 1ull << j, ++scale;
        } else if (cur < kSmall) {
          cur *= kBig, down |= 1ull << j, --scale;
        }
        wj = cur;
        if (scale > best_scale || (scale == best_scale && cur > best)) best_scale = scale, best = cur;
      }
    }
    w[j] = wj;
  }
  // bring every term onto the scale of the largest one: factors are exact powers of two, terms 2^-128 below the
  // maximum vanish
  const float inv_best = 1.0f / best;
  float z = 0.0f;
  int sc = 0;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    sc += int((up >> j) & 1ull) - int((down >> j) & 1ull);
    const int d = sc - best_scale;  // <= 0
    float v = w[j] * inv_best;
    v = (d == 0) ? v : (d == -1 ? v * kSmall : 0.0f);
    w[j] = v;
    z += v;
  }
  const float inv = 1.0f / z;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) acc[j] += w[j] * inv;
  return low;
}


// prob[j] = P(noise = low + j): the accumulate form on a zeroed array (tests)
template <int NMAX>
CB_HD int noise_window_pmf(float x, float mu, float lam, float alpha, int n, float (&prob)[NMAX]) {
#pragma unroll
  for (int j = 0; j < NMAX; ++j) prob[j] = 0.0f;
  return noise_window_pmf_accumulate<NMAX>(x, mu, lam, alpha, n, prob);
}

}  // namespace cb
