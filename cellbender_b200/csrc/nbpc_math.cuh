// fp32 device math for the NB (+) Poisson moment-matched likelihood and its gradients.
//
// Replaces the op chain of NegativeBinomialPoissonConvApprox.log_prob
// (reference: cellbender/remove_background/distributions/NegativeBinomialPoissonConvApprox.py:105-143,
//  formula :63-70) with ONE evaluation per element, kept in registers.
//
// The reference formula  L = lgamma(v+a) - lgamma(v+1) - lgamma(a) + a(log a - log(a+m)) + v(log m - log(a+m)),
// a = alpha_hat = (m/mu)^2 alpha, m = mu + lam, cancels catastrophically in fp32 when a is large
// (a reaches 1e10..1e25 whenever mu << lam; SURVEY.md finding 2).  We evaluate the SAME function
// in cancellation-free form so that the fp32 result tracks an fp64 evaluation of the reference:
//   w = m/a = mu^2/(m alpha)
//   L(v=0)  = -a log1p(w) = -m log1p(w)/w
//   L(v>0)  = T + L(0) + [v log m - lgamma(v+1)],   T = sum_{i<v} log1p((i-m)/(a+m))
//             (= lgamma(v+a) - lgamma(a) - v log(a+m) exactly, for integer v)
//   large v: T via a Stirling difference and the Poisson part in deviance form
//             v phi(m/v - 1) - 0.5 log(2 pi v) - S(v),  phi(u) = log1p(u) - u.
// Gradients, with A1 = a dL/da kept finite:
//   dL/dm = a (v-m) / (m (a+m)) = (v-m)/(m (1+w))
//   A1    = sum_{i<v} (m-i)/((a+i)(1+w)) + m h(w),   h(w) = 1/(1+w) - log1p(w)/w
//   dL/dmu = dL/dm - 2 A1 lam/(m mu);  dL/dlam = dL/dm + 2 A1/m;  dL/dalpha = A1/alpha
// Poisson branch (mu < 1e-5 lam): L = v log lam - lam - lgamma(v+1), dL/dlam = v/lam - 1, no grad to mu, alpha.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cb {

constexpr int kSmallCount = 16;  // counts <= this use exact integer recurrences

// The math below is __host__ __device__ so that tests/test_host_math.py can run the very same
// fp32 code on the CPU (no GPU in the build container) against the fp64 golden vectors.
#define CB_HD __host__ __device__ __forceinline__

// lgamma(v+1) = log(v!) for v = 0..16 (exact to fp32)
#define CB_LOGFACT_TABLE                                                                                   \
  {0.0f, 0.0f, 0.69314718056f, 1.79175946923f, 3.17805383035f, 4.78749174278f, 6.57925121201f,             \
   8.52516136107f, 10.6046029027f, 12.8018274801f, 15.1044125730f, 17.5023078459f, 19.9872144957f,         \
   22.5521638531f, 25.1912211827f, 27.8992713838f, 30.6718601061f}
__device__ __constant__ float kLogFactDev[kSmallCount + 1] = CB_LOGFACT_TABLE;
static const float kLogFactHost[kSmallCount + 1] = CB_LOGFACT_TABLE;
CB_HD float log_fact(int v) {
#ifdef __CUDA_ARCH__
  return kLogFactDev[v];
#else
  return kLogFactHost[v];
#endif
}

// phi(u) = log1p(u) - u, cancellation-free for small |u|
CB_HD float phi_f(float u) {
  if (fabsf(u) < 0.25f) {
    // u^2 * (-1/2 + u/3 - u^2/4 + ... - u^10/12)
    float p = -1.0f / 12.0f;
    p = fmaf(p, u, 1.0f / 11.0f);
    p = fmaf(p, u, -1.0f / 10.0f);
    p = fmaf(p, u, 1.0f / 9.0f);
    p = fmaf(p, u, -1.0f / 8.0f);
    p = fmaf(p, u, 1.0f / 7.0f);
    p = fmaf(p, u, -1.0f / 6.0f);
    p = fmaf(p, u, 1.0f / 5.0f);
    p = fmaf(p, u, -1.0f / 4.0f);
    p = fmaf(p, u, 1.0f / 3.0f);
    p = fmaf(p, u, -1.0f / 2.0f);
    return u * u * p;
  }
  return log1pf(u) - u;
}

// log(num/den) for num, den > 0; cancellation-free near 1 and safe when num << den
// (num_minus_den is passed separately so that it is not lost to rounding when den is huge)
CB_HD float log_ratio(float num, float den, float num_minus_den) {
  const float d = num_minus_den / den;
  return (fabsf(d) < 0.25f) ? log1pf(d) : logf(num / den);
}
// phi(t - 1) = log t - t + 1 for t > 0
CB_HD float phi_ratio(float t) { return (fabsf(t - 1.0f) < 0.25f) ? phi_f(t - 1.0f) : logf(t) - (t - 1.0f); }

// q(w) = log1p(w)/w  (-> 1 as w -> 0), w >= 0
CB_HD float log1p_over_x(float w) { return (w > 0.0f) ? log1pf(w) / w : 1.0f; }

// h(w) = 1/(1+w) - log1p(w)/w = sum_{k>=1} (-w)^k k/(k+1)
CB_HD float h_f(float w, float inv1pw, float q) {
  if (w < 0.25f) {
    float p = 11.0f / 12.0f;  // k = 11
    p = fmaf(p, -w, 10.0f / 11.0f);
    p = fmaf(p, -w, 9.0f / 10.0f);
    p = fmaf(p, -w, 8.0f / 9.0f);
    p = fmaf(p, -w, 7.0f / 8.0f);
    p = fmaf(p, -w, 6.0f / 7.0f);
    p = fmaf(p, -w, 5.0f / 6.0f);
    p = fmaf(p, -w, 4.0f / 5.0f);
    p = fmaf(p, -w, 3.0f / 4.0f);
    p = fmaf(p, -w, 2.0f / 3.0f);
    p = fmaf(p, -w, 1.0f / 2.0f);
    return -w * p;
  }
  return inv1pw - q;
}

// Stirling remainder S(z) = 1/(12 z) - 1/(360 z^3) + 1/(1260 z^5),  lgamma(z) = (z-.5) log z - z + .5 log(2 pi) + S(z)
CB_HD float stirling_S(float z) {
  const float r = 1.0f / z, r2 = r * r;
  return r * (1.0f / 12.0f + r2 * (-1.0f / 360.0f + r2 * (1.0f / 1260.0f)));
}
// S(z + d) - S(z), leading term in difference form (no cancellation when d << z)
CB_HD float stirling_S_diff(float z, float d) {
  const float z2 = z + d, r1 = 1.0f / z, r2 = 1.0f / z2;
  const float hi1 = r1 * r1 * r1 * (-1.0f / 360.0f + r1 * r1 * (1.0f / 1260.0f));
  const float hi2 = r2 * r2 * r2 * (-1.0f / 360.0f + r2 * r2 * (1.0f / 1260.0f));
  return -(1.0f / 12.0f) * d * r1 * r2 + (hi2 - hi1);
}
// psi(z) = log z + P(z),  P(z) = -1/(2z) - 1/(12 z^2) + 1/(120 z^4) - 1/(252 z^6);  returns P(z + d) - P(z)
CB_HD float digamma_P_diff(float z, float d) {
  const float z2 = z + d, r1 = 1.0f / z, r2 = 1.0f / z2;
  const float q1 = r1 * r1, q2 = r2 * r2;
  const float hi1 = q1 * q1 * (1.0f / 120.0f - q1 * (1.0f / 252.0f));
  const float hi2 = q2 * q2 * (1.0f / 120.0f - q2 * (1.0f / 252.0f));
  return 0.5f * d * r1 * r2 + (1.0f / 12.0f) * d * (z + z2) * q1 * q2 + (hi2 - hi1);
}

// log Poisson pmf  v log(rate) - rate - lgamma(v+1), v a non-negative integer-valued float
CB_HD float poisson_logpmf(float v, float rate) {
  if (v == 0.0f) return -rate;
  if (v <= float(kSmallCount)) return v * logf(rate) - rate - log_fact(int(v));
  // deviance form: v phi(rate/v - 1) - 0.5 log(2 pi v) - S(v)
  return v * phi_ratio(rate / v) - 0.5f * logf(6.283185307179586f * v) - stirling_S(v);
}

struct NbpcOut {
  float L, dmu, dlam, dalpha;
};

// ---- counts above kF64Count: the same cancellation-free decomposition evaluated in float64 -----------------------
// With counts in the hundreds the individual pieces (v phi(m/v - 1), a2 phi(u), v log ratios) reach 1e3 while L stays
// O(10): fp32 pieces leave 2-3e-5 of |L|.  Such elements are rare (a few stored counts per droplet), so they take the
// float64 pipe: same formulas, log1p-based phi (its rounding error scales with |u|, harmless in float64).
constexpr float kF64Count = 64.0f;
CB_HD double phi_d(double u) { return log1p(u) - u; }
CB_HD double log_ratio_d(double num, double den, double num_minus_den) {
  const double d = num_minus_den / den;
  return (fabs(d) < 0.25) ? log1p(d) : log(num / den);
}
CB_HD double stirling_S_d(double z) {
  const double r = 1.0 / z, r2 = r * r;
  return r * (1.0 / 12.0 + r2 * (-1.0 / 360.0 + r2 * (1.0 / 1260.0 - r2 * (1.0 / 1680.0))));
}
CB_HD double stirling_S_diff_d(double z, double d) {
  const double z2 = z + d, r1 = 1.0 / z, r2 = 1.0 / z2;
  const double q1 = r1 * r1, q2 = r2 * r2;
  const double hi1 = r1 * q1 * (-1.0 / 360.0 + q1 * (1.0 / 1260.0 - q1 * (1.0 / 1680.0)));
  const double hi2 = r2 * q2 * (-1.0 / 360.0 + q2 * (1.0 / 1260.0 - q2 * (1.0 / 1680.0)));
  return -(1.0 / 12.0) * d * r1 * r2 + (hi2 - hi1);
}
CB_HD double digamma_P_diff_d(double z, double d) {
  const double z2 = z + d, r1 = 1.0 / z, r2 = 1.0 / z2;
  const double q1 = r1 * r1, q2 = r2 * r2;
  const double hi1 = q1 * q1 * (1.0 / 120.0 - q1 * (1.0 / 252.0 - q1 * (1.0 / 240.0)));
  const double hi2 = q2 * q2 * (1.0 / 120.0 - q2 * (1.0 / 252.0 - q2 * (1.0 / 240.0)));
  return 0.5 * d * r1 * r2 + (1.0 / 12.0) * d * (z + z2) * q1 * q2 + (hi2 - hi1);
}

// (not inlined: a real call keeps its float64 registers out of the callers' hot fp32 paths)
template <bool GRAD>
__host__ __device__ __noinline__ NbpcOut nbpc_eval_large_f64(float vf, float muf, float alphaf, float lamf) {
  NbpcOut o;
  const double v = vf, mu = muf, alpha = alphaf, lam = lamf;
  const double m = mu + lam;
  const double w = (mu / m) * (mu / alpha);
  const double inv1pw = 1.0 / (1.0 + w);
  const double q = (w > 0.0) ? log1p(w) / w : 1.0;
  const double a = fmin(m / w, 1e300);
  const double apm = a + m;
  const int k = (a < 8.0) ? 8 : 0;  // shift small alpha_hat up so that the Stirling remainder applies
  double T = 0.0, S1 = 0.0;
  for (int i = 0; i < k; ++i) {
    const double fi = double(i);
    T += log_ratio_d(a + fi, apm, fi - m);
    if (GRAD) S1 += (m - fi) / ((a + fi) * (1.0 + w));
  }
  const double vr = v - double(k), a2 = a + double(k), m2 = m - double(k);
  const double u = vr / a2;
  T += a2 * phi_d(u) - 0.5 * log1p(u) + vr * log_ratio_d(a2 + vr, apm, vr - m2) + stirling_S_diff_d(a2, vr);
  const double t = m / v;
  const double pois = v * ((fabs(t - 1.0) < 0.25) ? phi_d(t - 1.0) : log(t) - (t - 1.0)) - 0.5 * log(6.283185307179586 * v)
                      - stirling_S_d(v);
  o.L = float(T - a * phi_d(w) + pois);
  o.dmu = 0.f, o.dlam = 0.f, o.dalpha = 0.f;
  if (GRAD) {
    S1 += a * (phi_d(u) + u * m2 / apm + digamma_P_diff_d(a2, vr));
    const double h = (w < 1e-4) ? -w * (0.5 - w * (2.0 / 3.0 - w * 0.75)) : inv1pw - q;
    const double dLdm = (v - m) / m * inv1pw;
    const double A1 = S1 + m * h;
    o.dmu = float(dLdm - 2.0 * A1 * (lam / m) / mu);
    o.dlam = float(dLdm + 2.0 * A1 / m);
    o.dalpha = float(A1 / alpha);
  }
  return o;
}

// One element of NBPCapprox.log_prob (+ gradients when GRAD).  mu, alpha, lam > 0; v >= 0 integer-valued.
template <bool GRAD>
CB_HD NbpcOut nbpc_eval(float v, float mu, float alpha, float lam) {
  NbpcOut o;
  o.dmu = 0.f, o.dlam = 0.f, o.dalpha = 0.f;
  if (mu < 1e-5f * lam) {  // reference: empty_indices = mu < 1e-5 * lam  -> Poisson(lam)
    o.L = poisson_logpmf(v, lam);
    if (GRAD) o.dlam = v / lam - 1.0f;
    return o;
  }
  if (v > kF64Count) return nbpc_eval_large_f64<GRAD>(v, mu, alpha, lam);
  const float m = mu + lam;
  const float w = (mu / m) * (mu / alpha);  // = m / alpha_hat
  const float inv1pw = 1.0f / (1.0f + w);
  const float q = log1p_over_x(w);
  const float L0 = -m * q;  // = -a log1p(m/a)
  float T = 0.f, S1 = 0.f;  // S1 = sum (m-i)/((a+i)(1+w))
  if (v > 0.0f) {
    const float a = fminf(m / w, 1e30f);  // alpha_hat (huge when mu << lam; clamp keeps inf*0 out)
    const float apm = a + m;
    int k;
    float vr = 0.f;  // counts handled by the Stirling remainder
    if (v <= float(kSmallCount)) {
      k = int(v);
    } else {
      k = (a < 8.0f) ? 8 : 0;
      vr = v - float(k);
    }
    for (int i = 0; i < k; ++i) {
      const float fi = float(i);
      T += log_ratio(a + fi, apm, fi - m);
      if (GRAD) S1 += (m - fi) / ((a + fi) * (1.0f + w));
    }
    if (vr > 0.0f) {
      // remaining counts: shift (a, m) -> (a + k, m - k); a' + m' = a + m
      const float a2 = a + float(k), m2 = m - float(k);
      const float u = vr / a2;
      T += a2 * phi_f(u) - 0.5f * log1pf(u) + vr * log_ratio(a2 + vr, apm, vr - m2) + stirling_S_diff(a2, vr);
      if (GRAD) {
        // a * [psi(a2+vr) - psi(a2) - vr/(a+m)] = a * [phi(u) + u m2/(a+m) + P(a2+vr) - P(a2)]
        S1 += a * (phi_f(u) + u * m2 / apm + digamma_P_diff(a2, vr));
      }
      o.L = T - a * phi_f(w) + poisson_logpmf(v, m);  // L0 + m = -a phi(w)
    } else {
      o.L = T + L0 + (v * logf(m) - log_fact(int(v)));
    }
  } else {
    o.L = L0;
  }
  if (GRAD) {
    const float dLdm = (v - m) / m * inv1pw;
    const float A1 = S1 + m * h_f(w, inv1pw, q);
    o.dmu = dLdm - 2.0f * A1 * (lam / m) / mu;
    o.dlam = dLdm + 2.0f * A1 / m;
    o.dalpha = A1 / alpha;
  }
  return o;
}

// x == 0 specialisation of nbpc_eval<true> (nbpc_math.cuh): L = -m q(w), q(w) = log1p(w)/w, w = mu^2 / (m alpha);
//   dL/dmu = -1/(1+w) - 2 h(w) lam/mu,  dL/dlam = 2 h(w) - 1/(1+w),  dL/dalpha = m h(w)/alpha,  h = 1/(1+w) - q.
// 33 538 genes of every droplet row take this path, so it is written with three reciprocals (one MUFU each) and, for
// w < 1/4 (mu << lam: the common case), degree-6 near-minimax polynomials on [0, 1/4] for q(w) and -h(w)/w
// (Chebyshev-node interpolants; max relative error 4.6e-10 and 6.3e-9 in exact arithmetic, i.e. below fp32 rounding;
// the 12-term Taylor series they replace cost twice the FMAs), straight-line with the reference's Poisson branch
// (mu < 1e-5 lam) applied as a select.
// Reciprocal of the dense x == 0 path: ONE rcp.approx.ftz (1 ulp; arguments are >= 1e-10, far from the flush range) on
// the device -- __fdividef(1, v) wraps the same instruction in a five-instruction range rescue, 10 % of the observation
// kernel's instructions in the r02i ncu capture -- exact on the host build.
CB_HD float rcp_fast(float v) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#else
  return 1.0f / v;
#endif
}

CB_HD NbpcOut nbpc_zero(float mu, float inv_alpha, float lam) {
  NbpcOut o;
  const float m = mu + lam;
  const float inv_m = rcp_fast(m);
  const float w = (mu * inv_m) * (mu * inv_alpha);
  const float inv1pw = rcp_fast(1.0f + w);
  float q, h;
  if (w < 0.25f) {
    float t = 7.0580221160e-02f;
    t = fmaf(t, w, -1.4531815925e-01f);
    t = fmaf(t, w, 1.9661888013e-01f);
    t = fmaf(t, w, -2.4971586912e-01f);
    t = fmaf(t, w, 3.3332176190e-01f);
    t = fmaf(t, w, -4.9999982126e-01f);
    q = fmaf(t, w, 1.0f);
    float u = 3.8702023389e-01f;
    u = fmaf(u, w, -7.1614602060e-01f);
    u = fmaf(u, w, 8.1124690281e-01f);
    u = fmaf(u, w, -7.9815522469e-01f);
    u = fmaf(u, w, 7.4992512791e-01f);
    u = fmaf(u, w, -6.6666551211e-01f);
    u = fmaf(u, w, 0.5f);
    h = -w * u;
  } else {
    q = log1pf(w) * (1.0f / w);
    h = inv1pw - q;
  }
  const bool poisson = mu < 1e-5f * lam;  // reference: Poisson(lam) where mu < 1e-5 lam
  o.L = poisson ? -lam : -m * q;
  o.dlam = poisson ? -1.0f : 2.0f * h - inv1pw;
  o.dmu = poisson ? 0.0f : -inv1pw - 2.0f * h * lam * rcp_fast(mu);
  o.dalpha = poisson ? 0.0f : m * h * inv_alpha;
  return o;
}


// x == 0 with mu equal to the bare safeguard constant mu_eps = 1e-10 (the y = 0 enumeration branch: no cell, model.py:
// 338-346 adds NBPC_MU_EPS_SAFEGAURD to a zero mean).  Then w = mu_eps^2 / (m alpha) <= 5e-11 / alpha is far below fp32
// resolution, and nbpc_zero collapses, to fp32 rounding, to
//   L = -(mu_eps + lam) (1 - w/2) = -(mu_eps + lam),   dL/dlam = -1 + w^2/3 = -1,
//   dL/dalpha = m h(w) / alpha = -mu_eps^2 / (2 alpha^2) (1 + O(w))
// (Poisson branch mu_eps < 1e-5 lam: L = -lam, dL/dlam = -1, dL/dalpha = 0, as in the reference.)  dL/dmu is not
// needed: mu_eps is a constant.  neg_half_mu2_inv_alpha2 = -mu_eps^2 / (2 alpha^2) is computed once per droplet row.
CB_HD NbpcOut nbpc_zero_mu_eps(float lam, float neg_half_mu2_inv_alpha2) {
  NbpcOut o;
  o.dmu = 0.f, o.dlam = -1.f;
  const bool poisson = 1e-10f < 1e-5f * lam;
  o.L = poisson ? -lam : -(1e-10f + lam);
  o.dalpha = poisson ? 0.f : neg_half_mu2_inv_alpha2;
  return o;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact NB(mu, alpha) (+) Poisson(lam) convolution by direct summation: NegativeBinomialPoissonConv.log_prob
// (reference distributions/NegativeBinomialPoissonConv.py:89-154; selected by consts.USE_EXACT_LOG_PROB, off by
// default).  L(v) = logsumexp_{k=0..max_poisson} [ Poisson(low + k; lam) + NB(v - low - k; mu, alpha) ],
// low = clamp(min(int(lam - max_poisson/2), int(v - max_poisson)), 0), terms with a negative NB count dropped.
// The reference's closed forms for v = 0 and v = 1 are this sum with one and two terms.
// One lgamma triple (in fp64) starts the two pmf recurrences; every further term costs two logs.  The gradients are
// the softmax-weighted term gradients, accumulated in the same pass (running-max rescaling).
// ---------------------------------------------------------------------------------------------------------------
CB_HD double digamma_f64(double x) {
  double r = 0.0;
  while (x < 6.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f / 132.0))));
}

template <bool GRAD>
CB_HD NbpcOut nbpc_exact_eval(float v, float mu, float alpha, float lam, int max_poisson) {
  NbpcOut o;
  o.dmu = 0.f, o.dlam = 0.f, o.dalpha = 0.f;
  int low = int(lam - 0.5f * float(max_poisson));
  const int low2 = int(v - float(max_poisson));
  low = low < low2 ? low : low2;
  low = low < 0 ? 0 : low;
  const int vi = int(v);
  if (low > vi) {  // every NB count would be negative: the reference's table is all -inf
    o.L = -INFINITY;
    return o;
  }
  const int n_terms = (vi - low < max_poisson ? vi - low : max_poisson) + 1;
  const double a = alpha, m = mu, l = lam;
  const double log_l = log(l), lar = log(a) - log(a + m), lmr = log(m) - log(a + m);
  double nb = double(vi - low), pois = double(low);
  // first term, directly
  double lp = pois * log_l - l - lgamma(pois + 1.0);
  double ln = lgamma(nb + a) - lgamma(nb + 1.0) - lgamma(a) + a * lar + nb * lmr;
  double dig = 0.0;  // psi(nb + alpha) - psi(alpha)
  if (GRAD) {
    if (nb <= 64.0) {
      for (int i = 0; i < int(nb); ++i) dig += 1.0 / (a + double(i));
    } else {
      dig = digamma_f64(nb + a) - digamma_f64(a);
    }
  }
  double mx = lp + ln, s = 1.0;
  double gl = 0.0, gm = 0.0, ga = 0.0;
  const double inv_apm = 1.0 / (a + m);
  if (GRAD) {
    gl = pois / l - 1.0;
    gm = nb / m - (a + nb) * inv_apm;
    ga = dig + lar + 1.0 - (a + nb) * inv_apm;
  }
  for (int k = 1; k < n_terms; ++k) {
    // Poisson(pois + 1) from Poisson(pois); NB(nb - 1) from NB(nb)
    lp += log_l - double(logf(float(pois + 1.0)));
    ln -= double(logf(float((nb - 1.0 + a) / nb))) + lmr;
    if (GRAD) dig -= 1.0 / (a + nb - 1.0);
    pois += 1.0;
    nb -= 1.0;
    const double t = lp + ln;
    double w;
    if (t > mx) {
      const double r = exp(mx - t);
      s = s * r + 1.0;
      if (GRAD) gl *= r, gm *= r, ga *= r;
      mx = t;
      w = 1.0;
    } else {
      w = exp(t - mx);
      s += w;
    }
    if (GRAD) {
      gl += w * (pois / l - 1.0);
      gm += w * (nb / m - (a + nb) * inv_apm);
      ga += w * (dig + lar + 1.0 - (a + nb) * inv_apm);
    }
  }
  o.L = float(mx + log(s));
  if (GRAD) {
    const double inv_s = 1.0 / s;
    o.dlam = float(gl * inv_s), o.dmu = float(gm * inv_s), o.dalpha = float(ga * inv_s);
  }
  return o;
}

}  // namespace cb
