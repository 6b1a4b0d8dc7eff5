// Elementwise drop-in for NegativeBinomialPoissonConvApprox.log_prob and its backward
// (reference: distributions/NegativeBinomialPoissonConvApprox.py:105-143; invoked at model.py:338-346).
// Logical shape [E, B, G]; each operand carries element strides (0 = broadcast), which is how
// value [B,G], mu/lam [2,B,G] and the scalar alpha arrive from pyro's enumeration.
#include "common.cuh"
#include "nbpc_math.cuh"

namespace cb {

struct Strides3 {
  long long e, b, g;
};

// EXACT = the full convolution (NegativeBinomialPoissonConv.log_prob); otherwise the moment-matched approximation
template <bool GRAD, bool EXACT>
__global__ void __launch_bounds__(256) nbpc_elementwise_kernel(
    const float* __restrict__ value, Strides3 sv, const float* __restrict__ mu, Strides3 sm,
    const float* __restrict__ alpha, Strides3 sa, const float* __restrict__ lam, Strides3 sl,
    const float* __restrict__ grad_out,  // [E,B,G] contiguous, GRAD only
    float* __restrict__ out,             // [E,B,G] log prob (forward) -- nullptr when GRAD
    float* __restrict__ dmu, float* __restrict__ dlam, float* __restrict__ dalpha,  // [E,B,G] contiguous
    int E, int B, int G, int* __restrict__ nan_flag, int max_poisson) {
  const long long total = (long long)E * B * G;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = int(idx % G);
    const long long t = idx / G;
    const int b = int(t % B), e = int(t / B);
    const float v = value[e * sv.e + b * sv.b + g * sv.g];
    const float m_ = mu[e * sm.e + b * sm.b + g * sm.g];
    const float a_ = alpha[e * sa.e + b * sa.b + g * sa.g];
    const float l_ = lam[e * sl.e + b * sl.b + g * sl.g];
    const NbpcOut o = EXACT ? nbpc_exact_eval<GRAD>(v, m_, a_, l_, max_poisson) : nbpc_eval<GRAD>(v, m_, a_, l_);
    if (!GRAD) {
      out[idx] = o.L;
      if (o.L != o.L) atomicOr(nan_flag, 1);
    } else {
      const float go = grad_out[idx];
      dmu[idx] = go * o.dmu;
      dlam[idx] = go * o.dlam;
      dalpha[idx] = go * o.dalpha;
    }
  }
}

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * cb::sm_count();
  return int(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace cb

extern "C" int cb_nbpc_approx_log_prob_fwd(const float* value, const long long* sv, const float* mu,
                                           const long long* sm, const float* alpha, const long long* sa,
                                           const float* lam, const long long* sl, float* out, int E, int B,
                                           int G, int* nan_flag, void* stream) {
  CB_REQUIRE(E > 0 && B > 0 && G > 0, "cb_nbpc_approx_log_prob_fwd: empty shape [%d,%d,%d]", E, B, G);
  const long long total = (long long)E * B * G;
  cb::nbpc_elementwise_kernel<false, false><<<cb::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      value, {sv[0], sv[1], sv[2]}, mu, {sm[0], sm[1], sm[2]}, alpha, {sa[0], sa[1], sa[2]}, lam,
      {sl[0], sl[1], sl[2]}, nullptr, out, nullptr, nullptr, nullptr, E, B, G, nan_flag, 0);
  return cb::check_launch("cb_nbpc_approx_log_prob_fwd");
}

extern "C" int cb_nbpc_approx_log_prob_bwd(const float* value, const long long* sv, const float* mu,
                                           const long long* sm, const float* alpha, const long long* sa,
                                           const float* lam, const long long* sl, const float* grad_out,
                                           float* dmu, float* dlam, float* dalpha, int E, int B, int G,
                                           void* stream) {
  CB_REQUIRE(E > 0 && B > 0 && G > 0, "cb_nbpc_approx_log_prob_bwd: empty shape [%d,%d,%d]", E, B, G);
  const long long total = (long long)E * B * G;
  cb::nbpc_elementwise_kernel<true, false><<<cb::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      value, {sv[0], sv[1], sv[2]}, mu, {sm[0], sm[1], sm[2]}, alpha, {sa[0], sa[1], sa[2]}, lam,
      {sl[0], sl[1], sl[2]}, grad_out, nullptr, dmu, dlam, dalpha, E, B, G, nullptr, 0);
  return cb::check_launch("cb_nbpc_approx_log_prob_bwd");
}

extern "C" int cb_nbpc_exact_log_prob_fwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                                          const float* alpha, const long long* sa, const float* lam, const long long* sl,
                                          float* out, int E, int B, int G, int max_poisson, int* nan_flag, void* stream) {
  CB_REQUIRE(E > 0 && B > 0 && G > 0 && max_poisson >= 0, "cb_nbpc_exact_log_prob_fwd: bad shape [%d,%d,%d]", E, B, G);
  const long long total = (long long)E * B * G;
  cb::nbpc_elementwise_kernel<false, true><<<cb::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      value, {sv[0], sv[1], sv[2]}, mu, {sm[0], sm[1], sm[2]}, alpha, {sa[0], sa[1], sa[2]}, lam,
      {sl[0], sl[1], sl[2]}, nullptr, out, nullptr, nullptr, nullptr, E, B, G, nan_flag, max_poisson);
  return cb::check_launch("cb_nbpc_exact_log_prob_fwd");
}

extern "C" int cb_nbpc_exact_log_prob_bwd(const float* value, const long long* sv, const float* mu, const long long* sm,
                                          const float* alpha, const long long* sa, const float* lam, const long long* sl,
                                          const float* grad_out, float* dmu, float* dlam, float* dalpha, int E, int B, int G,
                                          int max_poisson, void* stream) {
  CB_REQUIRE(E > 0 && B > 0 && G > 0 && max_poisson >= 0, "cb_nbpc_exact_log_prob_bwd: bad shape [%d,%d,%d]", E, B, G);
  const long long total = (long long)E * B * G;
  cb::nbpc_elementwise_kernel<true, true><<<cb::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      value, {sv[0], sv[1], sv[2]}, mu, {sm[0], sm[1], sm[2]}, alpha, {sa[0], sa[1], sa[2]}, lam,
      {sl[0], sl[1], sl[2]}, grad_out, nullptr, dmu, dlam, dalpha, E, B, G, nullptr, max_poisson);
  return cb::check_launch("cb_nbpc_exact_log_prob_bwd");
}
