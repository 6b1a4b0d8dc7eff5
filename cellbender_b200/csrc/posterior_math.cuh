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

// Fast device math, exactly ONE MUFU instruction each: lg2.approx.ftz / ex2.approx.ftz / rcp.approx.ftz (abs. / rel. error
// ~2^-22).  The .ftz forms matter: without them nvcc wraps every lg2 / ex2 / rcp in a denormal-range rescue (compare,
// select, two multiplies), which made the window kernel ALU-bound (ncu r02i: alu 51 %, xu 25 %).  Every argument here is
// far inside the normal range (ratios of counts and rates >= 1e-30 clamped to 3e37; exponents <= 0 whose flushed results
// lie dozens of orders of magnitude below the stored threshold).  The host build of the same source
// (tools/host_math_check.cu) uses the exact library functions.
CB_HD float pm_log2(float v) {
#ifdef __CUDA_ARCH__
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#else
  return log2f(v);
#endif
}
CB_HD float pm_exp2(float v) {
#ifdef __CUDA_ARCH__
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#else
  return exp2f(v);
#endif
}
CB_HD float pm_div(float a, float b) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
#else
  return a / b;
#endif
}

// Normalised pmf over the window, accumulated in place: acc[j] += P(noise = low + j), j < n (0 where low + j > x).
// mu, lam, alpha must already carry the reference's +1e-30 safeguards (posterior.py:708-710).
// Consecutive terms differ by the rational factor r(c) = (lam / s) (x - c) / ((c + 1) (alpha + x - c - 1)) with
// lam / s = lam (1 + alpha / mu); the relative log2-pmf is the running sum of log2 r, one rcp + one lg2 per term, and the
// terms are exponentiated with one ex2 each.  (The first version used logf / log1pf / expf and an IEEE division per
// term: 1 240 instructions per sample evaluation -- ncu r02k.)
// lam / s is clamped to 3e37: with mu -> 0 (+1e-30) all mass piles onto the last valid count, which the clamp reproduces
// to far below the stored threshold log p >= -10.
// The window has nlive = min(n, x - low + 1) live terms (c <= x); the rest are exact zeros.  Live-ness is a select, not a
// branch (r02i: the per-term `dead` flag cost 9 of 24 instructions per term); work beyond the live terms is skipped in
// blocks of four through ``live_bound``, an upper bound of nlive that the CALLER guarantees to be uniform over the warp
// (the window kernel passes the warp maximum of min(n, x + 1); the default evaluates every term).  Returns low.
template <int NMAX>
CB_HD int noise_window_pmf_accumulate(float x, float mu, float lam, float alpha, int n, float (&acc)[NMAX],
                                      int live_bound = NMAX) {
  const int low = noise_window_low(x, lam, n);
  const float lowf = float(low);
  const float kb = x - lowf;  // NB count at c = low: a non-negative integer (low <= max(x - n + 1, 0) <= x)
  const int nl = int(kb) + 1;
  const int nlive = nl < n ? nl : n;
  const float l_ratio = pm_log2(fminf(lam * (1.0f + pm_div(alpha, mu)), 3.0e37f));  // log2(lam / s)
  float t[NMAX];
  float tmax = 0.0f, run = 0.0f;
  t[0] = 0.0f;  // relative log2 prob of c = low
#pragma unroll
  for (int j0 = 0; j0 < NMAX; j0 += 4) {
    if (j0 < live_bound) {
#pragma unroll
      for (int j = (j0 == 0 ? 1 : j0); j < j0 + 4 && j < NMAX; ++j) {
        // ratio p(c + 1) / p(c) at c = low + j - 1: NB count k = x - c, alpha + k - 1 formed as alpha + (k - 1) (exact k - 1)
        const float k = kb - float(j - 1);
        const float r = pm_log2(pm_div(k, (lowf + float(j)) * (alpha + (kb - float(j)))));
        run = (j < nlive) ? run + (l_ratio + r) : -INFINITY;  // (a dead term's r may be NaN: never selected)
        tmax = fmaxf(tmax, run);
        t[j] = run;
      }
    } else {
#pragma unroll
      for (int j = j0; j < j0 + 4 && j < NMAX; ++j) t[j] = -INFINITY;
    }
  }
  float z = 0.0f;
#pragma unroll
  for (int j0 = 0; j0 < NMAX; j0 += 4) {
    if (j0 < live_bound) {
#pragma unroll
      for (int j = j0; j < j0 + 4 && j < NMAX; ++j) {
        t[j] = pm_exp2(t[j] - tmax);  // exp2(-inf) = 0 for the dead tail
        z += t[j];
      }
    }
  }
  const float inv = pm_div(1.0f, z);  // z >= 1: the largest term is exp2(0)
#pragma unroll
  for (int j0 = 0; j0 < NMAX; j0 += 4) {
    if (j0 < live_bound) {
#pragma unroll
      for (int j = j0; j < j0 + 4 && j < NMAX; ++j) acc[j] += t[j] * inv;
    }
  }
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
