// Host-side harness: runs the kernels' fp32 device math on the CPU (same source, __host__ path)
// so numerics can be validated in the GPU-less build container.  Test-only; not part of the product .so.
#include "../cellbender_b200/csrc/nbpc_math.cuh"
#include "../cellbender_b200/csrc/posterior_math.cuh"

extern "C" void host_nbpc_eval(int n, const float* v, const float* mu, const float* alpha, const float* lam,
                               float* L, float* dmu, float* dlam, float* dalpha) {
  for (int i = 0; i < n; ++i) {
    cb::NbpcOut o = cb::nbpc_eval<true>(v[i], mu[i], alpha[i], lam[i]);
    L[i] = o.L, dmu[i] = o.dmu, dlam[i] = o.dlam, dalpha[i] = o.dalpha;
  }
}

// prob: [n_entries, 32]; low: [n_entries]
extern "C" void host_noise_window_pmf(int n_entries, const float* x, const float* mu, const float* lam,
                                      const float* alpha, int n, float* prob, int* low) {
  for (int i = 0; i < n_entries; ++i) {
    float p[32];
    low[i] = cb::noise_window_pmf<32>(x[i], mu[i], lam[i], alpha[i], n, p);
    for (int j = 0; j < 32; ++j) prob[i * 32 + j] = p[j];
  }
}

// the x == 0 specialisation used by the dense pass of the fused observation kernel
extern "C" void host_nbpc_zero(int n, const float* mu, const float* alpha, const float* lam, float* L, float* dmu,
                               float* dlam, float* dalpha) {
  for (int i = 0; i < n; ++i) {
    cb::NbpcOut o = cb::nbpc_zero(mu[i], 1.0f / alpha[i], lam[i]);
    L[i] = o.L, dmu[i] = o.dmu, dlam[i] = o.dlam, dalpha[i] = o.dalpha;
  }
}

// the exact NB (+) Poisson convolution (NegativeBinomialPoissonConv.log_prob)
extern "C" void host_nbpc_exact(int n, const float* v, const float* mu, const float* alpha, const float* lam, int max_poisson,
                                float* L, float* dmu, float* dlam, float* dalpha) {
  for (int i = 0; i < n; ++i) {
    cb::NbpcOut o = cb::nbpc_exact_eval<true>(v[i], mu[i], alpha[i], lam[i], max_poisson);
    L[i] = o.L, dmu[i] = o.dmu, dlam[i] = o.dlam, dalpha[i] = o.dalpha;
  }
}

// ... and its y = 0 collapse (mu = the 1e-10 safeguard constant)
extern "C" void host_nbpc_zero_mu_eps(int n, const float* alpha, const float* lam, float* L, float* dlam, float* dalpha) {
  for (int i = 0; i < n; ++i) {
    const float ia = 1.0f / alpha[i];
    cb::NbpcOut o = cb::nbpc_zero_mu_eps(lam[i], -0.5f * 1e-10f * 1e-10f * ia * ia);
    L[i] = o.L, dlam[i] = o.dlam, dalpha[i] = o.dalpha;
  }
}
