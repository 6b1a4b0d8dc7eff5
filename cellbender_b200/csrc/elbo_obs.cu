// K5+K6+K7: the observation terms of the per-minibatch ELBO, forward AND backward, in one launch.
//
// Replaces (reference cellbender/remove_background, per svi.step):
//   Decoder softmax                                   vae/decoder.py:46-47
//   calculate_mu / calculate_lambda, y enumerated     model.py:25-85, 308-320   ([2,B,G] temporaries)
//   NBPCapprox(...).to_event(1).log_prob(x)           model.py:338-346 -> distributions/...ConvApprox.py:105-143
//   Poisson "obs_empty" regulariser                   model.py:374-398
//   and the autograd backward of all of the above (approx. 60 eager kernels + their saved tensors).
//
// Inputs per droplet row b (one CTA per row):
//   logits[b, :]        decoder pre-softmax output (chi = softmax(logits))
//   CSR row             the observed counts x[b, :] (x == 0 everywhere else)
//   coef[b, 0..6]       a_mu1, cA1, cB1, cA0, cB0, cAE, cBE  -- the rate mixtures in factored form:
//                         mu_y1  = a_mu1 * chi      + 1e-10      (mu_y0 = 1e-10)
//                         lam_y  = cAy * chi_ambient + cBy * chi_bar + 1e-10     (y = 1, 0)
//                         lam_E  = cAE * chi_ambient + cBE * chi_bar + 1e-10     (empty-droplet regulariser)
//                       (model.py:49-54,78-80 for every model_type reduce to this form; the host computes
//                        the [B]-sized coefficients, so the kernel is model-type agnostic)
//   w[b, 0..2]          d(loss)/d(L_y1), d(loss)/d(L_y0), d(loss)/d(L_E): the enumeration (Dice) weights
//                       times the loss sign/scale, known before the likelihood is evaluated.
// Outputs:
//   sums[b, 0..11]      L_y1, L_y0, L_E, S_mu_chi1, S_lam_a1, S_lam_b1, dalpha1, S_lam_a0, S_lam_b0, dalpha0,
//                       S_lam_aE, S_lam_bE   (S_* = sum_g dL/drate * profile: the exact gradients of L wrt coef)
//   dlogits[b, :]       d(loss)/d(logits) (softmax backward fused)
//   qamb[b, :]          row b's contribution to d(loss)/d(chi_ambient); column-summed by cb_colsum_rows
//
// Three phases per CTA: (A) row logsumexp; (B) dense pass with the x == 0 closed forms + sparse pass over
// the row's non-zeros with the full count formulas (a shared-memory bitmap keeps the two disjoint);
// (C) softmax backward.  Phases B and C re-read the row from L2/L1, not HBM.
#include "common.cuh"
#include "nbpc_math.cuh"

namespace cb {

constexpr int kObsThreads = 512;
constexpr int kNumSums = 12;

struct RowCoef {
  float a_mu1, cA1, cB1, cA0, cB0, cAE, cBE;
  float w1, w0, wE;
  float alpha;
};

// contributions of one (b, g) element; adds into acc[], returns p = dL1/dmu1 * chi and q (dchi_ambient contribution).
// MAXP < 0: NegativeBinomialPoissonConvApprox (the reference's default); MAXP >= 0: the exact convolution
// NegativeBinomialPoissonConv(max_poisson = MAXP) (model.py:336-357 with consts.USE_EXACT_LOG_PROB).
template <bool EXACT>
__device__ __forceinline__ void obs_element(float v, float chi, float ca, float cb_, const RowCoef& c, int max_poisson,
                                            float (&acc)[kNumSums], float& p_out, float& q_out) {
  const float mu1 = fmaf(c.a_mu1, chi, 1e-10f);
  const float lam1 = fmaf(c.cA1, ca, fmaf(c.cB1, cb_, 1e-10f));
  const float lam0 = fmaf(c.cA0, ca, fmaf(c.cB0, cb_, 1e-10f));
  const float lamE = fmaf(c.cAE, ca, fmaf(c.cBE, cb_, 1e-10f));
  const NbpcOut o1 = EXACT ? nbpc_exact_eval<true>(v, mu1, c.alpha, lam1, max_poisson) : nbpc_eval<true>(v, mu1, c.alpha, lam1);
  const NbpcOut o0 = EXACT ? nbpc_exact_eval<true>(v, 1e-10f, c.alpha, lam0, max_poisson) : nbpc_eval<true>(v, 1e-10f, c.alpha, lam0);
  const float LE = poisson_logpmf(v, lamE);
  const float dE = v / lamE - 1.0f;
  acc[0] += o1.L;
  acc[1] += o0.L;
  acc[2] += LE;
  acc[3] += o1.dmu * chi;
  acc[4] += o1.dlam * ca;
  acc[5] += o1.dlam * cb_;
  acc[6] += o1.dalpha;
  acc[7] += o0.dlam * ca;
  acc[8] += o0.dlam * cb_;
  acc[9] += o0.dalpha;
  acc[10] += dE * ca;
  acc[11] += dE * cb_;
  p_out = o1.dmu * chi;
  q_out = c.w1 * c.cA1 * o1.dlam + c.w0 * c.cA0 * o0.dlam + c.wE * c.cAE * dE;
}

// x == 0 under the exact convolution: the single term Poisson(0; lam) NB(0; mu, alpha) (...Conv.py:104-108)
//   L = -lam + alpha log(alpha / (alpha + mu)),  dL/dmu = -alpha / (alpha + mu),  dL/dlam = -1,
//   dL/dalpha = log(alpha / (alpha + mu)) + mu / (alpha + mu)
__device__ __forceinline__ NbpcOut nbpc_exact_zero(float mu, float alpha, float inv_alpha, float lam) {
  NbpcOut o;
  const float r = mu * inv_alpha;
  const float lar = -log1pf(r);
  const float f = 1.0f / (1.0f + r);  // alpha / (alpha + mu)
  o.L = fmaf(alpha, lar, -lam);
  o.dmu = -f;
  o.dlam = -1.0f;
  o.dalpha = lar + r * f;
  return o;
}

// contributions of one gene WITHOUT a stored count (x == 0)
template <bool EXACT>
__device__ __forceinline__ void obs_zero_element(float chi, float ca, float cb_, const RowCoef& c, float inv_alpha,
                                                 float (&acc)[kNumSums], float& p_out, float& q_out) {
  const float mu1 = fmaf(c.a_mu1, chi, 1e-10f);
  const float lam1 = fmaf(c.cA1, ca, fmaf(c.cB1, cb_, 1e-10f));
  const float lam0 = fmaf(c.cA0, ca, fmaf(c.cB0, cb_, 1e-10f));
  const float lamE = fmaf(c.cAE, ca, fmaf(c.cBE, cb_, 1e-10f));
  const float neg_half_mu2_inv_alpha2 = -0.5f * 1e-10f * 1e-10f * inv_alpha * inv_alpha;
  const NbpcOut o1 = EXACT ? nbpc_exact_zero(mu1, c.alpha, inv_alpha, lam1) : nbpc_zero(mu1, inv_alpha, lam1);
  // exact, mu = 1e-10: alpha log(alpha / (alpha + mu)) = -mu to fp32, dL/dalpha = -(mu / alpha)^2 / 2 (no Poisson branch)
  const NbpcOut o0 = EXACT ? NbpcOut{-(1e-10f + lam0), 0.f, -1.0f, neg_half_mu2_inv_alpha2} : nbpc_zero_mu_eps(lam0, neg_half_mu2_inv_alpha2);
  acc[0] += o1.L;
  acc[1] += o0.L;
  acc[2] -= lamE;  // Poisson(0; lamE)
  acc[3] += o1.dmu * chi;
  acc[4] += o1.dlam * ca;
  acc[5] += o1.dlam * cb_;
  acc[6] += o1.dalpha;
  acc[7] += o0.dlam * ca;
  acc[8] += o0.dlam * cb_;
  acc[9] += o0.dalpha;
  acc[10] -= ca;
  acc[11] -= cb_;
  p_out = o1.dmu * chi;
  q_out = c.w1 * c.cA1 * o1.dlam + c.w0 * c.cA0 * o0.dlam - c.wE * c.cAE;
}

template <bool EXACT>
__global__ void __launch_bounds__(kObsThreads, 2) elbo_obs_kernel(
    const long long* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
    const long long* __restrict__ row_ids, int G, const float* __restrict__ logits, long long ldl,
    const float* __restrict__ chi_amb, const float* __restrict__ chi_bar, const float* __restrict__ coef,
    const float* __restrict__ alpha, const float* __restrict__ w, float* __restrict__ sums,
    float* __restrict__ dlogits, float* __restrict__ qamb, long long ldq, float* __restrict__ lse_out,
    int* __restrict__ nan_flag, int max_poisson) {
  extern __shared__ unsigned int bitmap[];  // (G + 31) / 32 words
  __shared__ float red[32];
  __shared__ double dred[kNumSums * 32];

  const int b = blockIdx.x;
  const long long r = row_ids[b];
  const long long s = indptr[r], e = indptr[r + 1];
  const float* lrow = logits + (long long)b * ldl;
  float* drow = dlogits + (long long)b * ldl;
  float* qrow = qamb + (long long)b * ldq;
  const int nwords = (G + 31) >> 5;

  RowCoef c;
  c.a_mu1 = coef[b * 7 + 0], c.cA1 = coef[b * 7 + 1], c.cB1 = coef[b * 7 + 2], c.cA0 = coef[b * 7 + 3];
  c.cB0 = coef[b * 7 + 4], c.cAE = coef[b * 7 + 5], c.cBE = coef[b * 7 + 6];
  c.w1 = w[b * 3 + 0], c.w0 = w[b * 3 + 1], c.wE = w[b * 3 + 2];
  c.alpha = alpha[0];
  const float inv_alpha = 1.0f / c.alpha;

  // ---- phase 0: bitmap of genes with a stored count ----
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0u;
  __syncthreads();
  for (long long j = s + threadIdx.x; j < e; j += blockDim.x) {
    const int g = indices[j];
    atomicOr(&bitmap[g >> 5], 1u << (g & 31));
  }

  // ---- phase A: row logsumexp ----
  const int G4 = G >> 2;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < G4; i += blockDim.x) {
    const float4 l = *reinterpret_cast<const float4*>(lrow + 4 * i);
    mx = fmaxf(fmaxf(mx, fmaxf(l.x, l.y)), fmaxf(l.z, l.w));
  }
  for (int g = 4 * G4 + threadIdx.x; g < G; g += blockDim.x) mx = fmaxf(mx, lrow[g]);
  mx = block_max(mx, red);
  float se[1] = {0.f};
  for (int i = threadIdx.x; i < G4; i += blockDim.x) {
    const float4 l = *reinterpret_cast<const float4*>(lrow + 4 * i);
    se[0] += __expf(l.x - mx) + __expf(l.y - mx) + __expf(l.z - mx) + __expf(l.w - mx);
  }
  for (int g = 4 * G4 + threadIdx.x; g < G; g += blockDim.x) se[0] += __expf(lrow[g] - mx);
  block_sum<float, 1>(se, red);  // (contains the __syncthreads that also publishes the bitmap)
  const float lse = mx + logf(se[0]);
  if (threadIdx.x == 0 && lse_out != nullptr) lse_out[b] = lse;

  // ---- phase B (dense): every gene without a stored count, x == 0 ----
  float acc[kNumSums];
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) acc[i] = 0.f;
  for (int i = threadIdx.x; i < G4; i += blockDim.x) {
    const int g0 = 4 * i;
    const float4 l = *reinterpret_cast<const float4*>(lrow + g0);
    const float4 a4 = *reinterpret_cast<const float4*>(chi_amb + g0);
    const float4 b4 = *reinterpret_cast<const float4*>(chi_bar + g0);
    const unsigned int bits = (bitmap[g0 >> 5] >> (g0 & 31)) & 0xFu;
    const float ll[4] = {l.x, l.y, l.z, l.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float pp[4], qq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pp[k] = 0.f, qq[k] = 0.f;
      if (!((bits >> k) & 1u)) obs_zero_element<EXACT>(__expf(ll[k] - lse), aa[k], bb[k], c, inv_alpha, acc, pp[k], qq[k]);
    }
    *reinterpret_cast<float4*>(drow + g0) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(qrow + g0) = make_float4(qq[0], qq[1], qq[2], qq[3]);
  }
  for (int g = 4 * G4 + threadIdx.x; g < G; g += blockDim.x) {
    float p = 0.f, q = 0.f;
    if (!((bitmap[g >> 5] >> (g & 31)) & 1u))
      obs_zero_element<EXACT>(__expf(lrow[g] - lse), chi_amb[g], chi_bar[g], c, inv_alpha, acc, p, q);
    drow[g] = p;
    qrow[g] = q;
  }
  __syncthreads();  // dense stores to drow/qrow precede the sparse overwrites below

  // ---- phase B (sparse): the stored counts ----
  for (long long j = s + threadIdx.x; j < e; j += blockDim.x) {
    const int g = indices[j];
    float p, q;
    obs_element<EXACT>(data[j], expf(lrow[g] - lse), chi_amb[g], chi_bar[g], c, max_poisson, acc, p, q);
    drow[g] = p;
    qrow[g] = q;
  }

  // ---- row reductions (double, fixed tree -> deterministic) ----
  double dacc[kNumSums];
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) dacc[i] = double(acc[i]);
  block_sum<double, kNumSums>(dacc, dred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kNumSums; ++i) sums[b * kNumSums + i] = float(dacc[i]);
    const float chk = float(dacc[0] + dacc[1] + dacc[2]);
    if (chk != chk) atomicOr(nan_flag, 1);
  }
  const float s_mu_chi = float(dacc[3]);
  const float scale = c.w1 * c.a_mu1;
  __syncthreads();  // sparse stores visible before phase C reads drow

  // ---- phase C: softmax backward  dlogit = w1 a_mu1 (p - chi * sum_g p) ----
  for (int i = threadIdx.x; i < G4; i += blockDim.x) {
    const int g0 = 4 * i;
    const float4 l = *reinterpret_cast<const float4*>(lrow + g0);
    float4 p = *reinterpret_cast<float4*>(drow + g0);
    p.x = scale * (p.x - __expf(l.x - lse) * s_mu_chi);
    p.y = scale * (p.y - __expf(l.y - lse) * s_mu_chi);
    p.z = scale * (p.z - __expf(l.z - lse) * s_mu_chi);
    p.w = scale * (p.w - __expf(l.w - lse) * s_mu_chi);
    *reinterpret_cast<float4*>(drow + g0) = p;
  }
  for (int g = 4 * G4 + threadIdx.x; g < G; g += blockDim.x)
    drow[g] = scale * (drow[g] - __expf(lrow[g] - lse) * s_mu_chi);
}

// Deterministic column sum of a [R, ld] matrix: out[g] = sum_r in[r, g].  A CTA owns 32 adjacent columns; its 256
// threads are 8 row groups x 32 columns (coalesced 128-byte row segments, 32 independent loads in flight per thread);
// the row groups are combined in shared memory in a fixed order.
constexpr int kColsumCols = 32, kColsumGroups = 8, kColsumUnroll = 32;
__global__ void __launch_bounds__(kColsumCols* kColsumGroups) colsum_rows_kernel(const float* __restrict__ in, int R, int G,
                                                                                long long ld, float* __restrict__ out) {
  __shared__ float s[kColsumGroups][kColsumCols + 1];
  const int c = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int g = blockIdx.x * kColsumCols + c;
  const bool ok = g < G;
  float acc = 0.f;
  for (int r0 = grp; r0 < R; r0 += kColsumGroups * kColsumUnroll) {
    float v[kColsumUnroll];
#pragma unroll
    for (int i = 0; i < kColsumUnroll; ++i) {
      const int r = r0 + i * kColsumGroups;
      v[i] = (ok && r < R) ? in[(long long)r * ld + g] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kColsumUnroll; ++i) acc += v[i];
  }
  s[grp][c] = acc;
  __syncthreads();
  if (grp == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kColsumGroups; ++i) t += s[i][c];
    out[g] = t;
  }
}

}  // namespace cb

static int elbo_obs_launch(const long long* indptr, const int* indices, const float* data, const long long* row_ids, int n_rows,
                           int n_genes, const float* logits, long long ldl, const float* chi_ambient, const float* chi_bar,
                           const float* coef, const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                           long long ldq, float* lse_out, int* nan_flag, int max_poisson, void* stream) {
  if (n_rows == 0) return 0;
  CB_REQUIRE(n_rows > 0 && n_genes > 0, "cb_elbo_obs_fwd_bwd: bad shape [%d, %d]", n_rows, n_genes);
  CB_REQUIRE((ldl & 3) == 0 && (ldq & 3) == 0 && ldl >= n_genes && ldq >= n_genes,
             "cb_elbo_obs_fwd_bwd: ldl=%lld / ldq=%lld must be >= n_genes and multiples of 4", ldl, ldq);
  for (const void* p : {(const void*)logits, (const void*)dlogits, (const void*)qamb, (const void*)chi_ambient,
                        (const void*)chi_bar})
    CB_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "cb_elbo_obs_fwd_bwd: pointer not 16-byte aligned");
  const size_t smem = size_t((n_genes + 31) / 32) * sizeof(unsigned int);
  CB_REQUIRE(smem <= 200 * 1024, "cb_elbo_obs_fwd_bwd: n_genes=%d too large for the shared-memory bitmap", n_genes);
  auto kern = max_poisson >= 0 ? cb::elbo_obs_kernel<true> : cb::elbo_obs_kernel<false>;
  if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  kern<<<n_rows, cb::kObsThreads, smem, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_genes, logits, ldl, chi_ambient,
                                                                chi_bar, coef, alpha, w, sums, dlogits, qamb, ldq, lse_out,
                                                                nan_flag, max_poisson);
  return cb::check_launch("cb_elbo_obs_fwd_bwd");
}

extern "C" int cb_elbo_obs_fwd_bwd(const long long* indptr, const int* indices, const float* data,
                                   const long long* row_ids, int n_rows, int n_genes, const float* logits,
                                   long long ldl, const float* chi_ambient, const float* chi_bar, const float* coef,
                                   const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                                   long long ldq, float* lse_out, int* nan_flag, void* stream) {
  return elbo_obs_launch(indptr, indices, data, row_ids, n_rows, n_genes, logits, ldl, chi_ambient, chi_bar, coef, alpha, w, sums,
                         dlogits, qamb, ldq, lse_out, nan_flag, -1, stream);
}

// The same with the exact NB (+) Poisson convolution as the likelihood (NegativeBinomialPoissonConv, max_poisson terms).
extern "C" int cb_elbo_obs_exact_fwd_bwd(const long long* indptr, const int* indices, const float* data,
                                         const long long* row_ids, int n_rows, int n_genes, const float* logits,
                                         long long ldl, const float* chi_ambient, const float* chi_bar, const float* coef,
                                         const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                                         long long ldq, float* lse_out, int* nan_flag, int max_poisson, void* stream) {
  CB_REQUIRE(max_poisson >= 0 && max_poisson <= 1000, "cb_elbo_obs_exact_fwd_bwd: max_poisson=%d not in [0, 1000]", max_poisson);
  return elbo_obs_launch(indptr, indices, data, row_ids, n_rows, n_genes, logits, ldl, chi_ambient, chi_bar, coef, alpha, w, sums,
                         dlogits, qamb, ldq, lse_out, nan_flag, max_poisson, stream);
}

extern "C" int cb_colsum_rows(const float* in, int n_rows, int n_cols, long long ld, float* out, void* stream) {
  CB_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "cb_colsum_rows: bad shape [%d, %d] ld=%lld", n_rows, n_cols, ld);
  cb::colsum_rows_kernel<<<(n_cols + cb::kColsumCols - 1) / cb::kColsumCols, cb::kColsumCols * cb::kColsumGroups, 0,
                           (cudaStream_t)stream>>>(in, n_rows, n_cols, ld, out);
  return cb::check_launch("cb_colsum_rows");
}
