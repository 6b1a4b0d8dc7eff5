// K5+K6+K7: the observation terms of the per-minibatch ELBO, forward AND backward.
//
// Replaces (reference cellbender/remove_background, per svi.step):
//   Decoder softmax                                   vae/decoder.py:46-47
//   calculate_mu / calculate_lambda, y enumerated     model.py:25-85, 308-320   ([2,B,G] temporaries)
//   NBPCapprox(...).to_event(1).log_prob(x)           model.py:338-346 -> distributions/...ConvApprox.py:105-143
//   Poisson "obs_empty" regulariser                   model.py:374-398
//   and the autograd backward of all of the above (approx. 60 eager kernels + their saved tensors).
//
// Inputs per droplet row b:
//   logits[b, :]        decoder pre-softmax output (chi = softmax(logits))
//   CSR row             the observed counts x[b, :] (x == 0 everywhere else)
//   coef[b, 0..6]       a_mu1, cA1, cB1, cA0, cB0, cAE, cBE  -- the rate mixtures in factored form:
//                         mu_y1  = a_mu1 * chi      + 1e-10      (mu_y0 = 1e-10)
//                         lam_y  = cAy * chi_ambient + cBy * chi_bar + 1e-10     (y = 1, 0)
//                         lam_E  = cAE * chi_ambient + cBE * chi_bar + 1e-10     (empty-droplet regulariser)
//                       (model.py:49-54,78-80 for every model_type reduce to this form; the host computes
//                        the [B]-sized coefficients, so the kernels are model-type agnostic)
//   w[b, 0..2]          d(loss)/d(L_y1), d(loss)/d(L_y0), d(loss)/d(L_E): the enumeration (Dice) weights
//                       times the loss sign/scale, known before the likelihood is evaluated.
// Outputs:
//   sums[b, 0..11]      L_y1, L_y0, L_E, S_mu_chi1, S_lam_a1, S_lam_b1, dalpha1, S_lam_a0, S_lam_b0, dalpha0,
//                       S_lam_aE, S_lam_bE   (S_* = sum_g dL/drate * profile: the exact gradients of L wrt coef)
//   dlogits[b, :]       d(loss)/d(logits) (softmax backward fused)
//   qamb[b, :]          row b's contribution to d(loss)/d(chi_ambient); column-summed by cb_colsum_rows
//
// Work decomposition (r02j).  Until r02i one CTA owned a whole droplet row: 255 CTAs on 296 slots, and a cell row (its
// stored counts take the full count formulas) costs twice an empty one -- the ncu capture showed the average SM busy for
// 61 % of the kernel's duration.  Now a row is cut into kObsChunks gene ranges and the three phases are three launches
// over (chunk, row) tiles with nothing but the stream order between them:
//   obs_row_lse_kernel       per tile (max, sum exp) of the logits            -> scratch
//   obs_main_kernel          row logsumexp from the tiles' pairs; dense pass with the x == 0 closed form; sparse pass
//                            over the tile's stored counts with the full count formulas (a shared-memory bitmap keeps
//                            the two disjoint); p = dL1/dmu1 * chi into dlogits, q into qamb; 12 partial row sums in
//                            float64 -> scratch
//   obs_softmax_bwd_kernel   row sums = fixed-order sum of the tiles' partials; dlogits = w1 a_mu1 (p - chi sum_g p)
// All reductions run in a fixed order (deterministic).  In the dense pass the y = 0 and empty-droplet branches are
// linear in (chi_ambient, chi_bar) where x == 0 (L = -lam, dL/dlam = -1), so they are accumulated as two plain sums of
// the profiles over the tile's zero-count genes and expanded once per thread, not evaluated per gene.
#include "common.cuh"
#include "nbpc_math.cuh"

namespace cb {

// tile shape of the three launches; overridable at compile time for tools/obs_variants.py (A/B builds)
#ifndef CB_OBS_CHUNKS
#define CB_OBS_CHUNKS 4
#endif
#ifndef CB_OBS_MIN_CTAS
#define CB_OBS_MIN_CTAS 4  // 64 registers: 592 of the 1020 tiles resident (r02k A/B: 86.4 us against 92.2 us with 3, same bits)
#endif
constexpr int kObsThreads = 256;
constexpr int kObsChunks = CB_OBS_CHUNKS;  // gene ranges per droplet row
constexpr int kNumSums = 12;
constexpr int kObsScratchPerTile = 2 * kNumSums + 2;  // floats: 12 float64 partial sums + (max, sum exp)

// genes per tile: a multiple of 128, so that bitmap words and float4 groups never straddle two tiles
__host__ __device__ inline int obs_chunk_len(int G) { return ((G + kObsChunks * 128 - 1) / (kObsChunks * 128)) * 128; }

struct RowCoef {
  float a_mu1, cA1, cB1, cA0, cB0, cAE, cBE;
  float w1, w0, wE;
  float alpha;
};

__device__ __forceinline__ float ex2_fast(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// exp(v) as one FMUL + one ex2.approx.ftz: relative error 2^-22; v <= 0 here (softmax terms), a flushed result is < 1e-38
__device__ __forceinline__ float exp_fast(float v) { return ex2_fast(v * 1.4426950408889634f); }

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

// Dense-pass accumulators of one thread: the y = 1 branch of the genes WITHOUT a stored count, and the plain sums of
// the two profiles over those genes (from which the y = 0 and empty-droplet branches follow, see obs_dense_expand).
struct DenseAcc {
  float L1, s_mu_chi, s_lam_a, s_lam_b, dalpha, sum_a, sum_b;
};

// one gene without a stored count (x == 0): the y = 1 branch; q_lin = -(w0 cA0 + wE cAE), k1 = w1 cA1 (per row)
template <bool EXACT>
__device__ __forceinline__ void obs_zero_element(float chi, float ca, float cb_, const RowCoef& c, float inv_alpha, float k1,
                                                 float q_lin, DenseAcc& a, float& p_out, float& q_out) {
  const float mu1 = fmaf(c.a_mu1, chi, 1e-10f);
  const float lam1 = fmaf(c.cA1, ca, fmaf(c.cB1, cb_, 1e-10f));
  const NbpcOut o1 = EXACT ? nbpc_exact_zero(mu1, c.alpha, inv_alpha, lam1) : nbpc_zero(mu1, inv_alpha, lam1);
  const float p = o1.dmu * chi;
  a.L1 += o1.L;
  a.s_mu_chi += p;
  a.s_lam_a = fmaf(o1.dlam, ca, a.s_lam_a);
  a.s_lam_b = fmaf(o1.dlam, cb_, a.s_lam_b);
  a.dalpha += o1.dalpha;
  a.sum_a += ca;
  a.sum_b += cb_;
  p_out = p;
  q_out = fmaf(k1, o1.dlam, q_lin);
}

// The y = 0 branch (mu = 1e-10) and the empty-droplet Poisson regulariser at x == 0, summed over the genes a thread
// visited in the dense pass, from the two profile sums:
//   L_y0 = -sum (cA0 ca + cB0 cb + 1e-10),   dL_y0/dlam = -1   =>  S_lam_a0 = -sum ca, S_lam_b0 = -sum cb
//   L_E  = -sum (cAE ca + cBE cb + 1e-10),   dL_E/dlam  = -1   =>  S_lam_aE = -sum ca, S_lam_bE = -sum cb
// The n_zero * 1e-10 constants are added once per tile (obs_main_kernel).  Approximate likelihood: where
// mu_eps = 1e-10 >= 1e-5 lam (no Poisson branch, lam <= 1e-5) the closed form differs from -lam by another 1e-10 in L
// and by -mu_eps^2 / (2 alpha^2) ~ 1e-21 in dL/dalpha per gene: at most 3e-6 and 2e-16 per droplet against sums of
// 1e2..1e4 and O(1), below fp32 resolution of the sums -- dropped.  The exact convolution's x == 0 term has
// dL/dalpha = -mu_eps^2 / (2 alpha^2) for every gene: added with the constants.
__device__ __forceinline__ void obs_dense_expand(const DenseAcc& a, const RowCoef& c, float (&acc)[kNumSums]) {
  acc[0] = a.L1;
  acc[1] = -fmaf(c.cA0, a.sum_a, c.cB0 * a.sum_b);
  acc[2] = -fmaf(c.cAE, a.sum_a, c.cBE * a.sum_b);
  acc[3] = a.s_mu_chi;
  acc[4] = a.s_lam_a;
  acc[5] = a.s_lam_b;
  acc[6] = a.dalpha;
  acc[7] = -a.sum_a;
  acc[8] = -a.sum_b;
  acc[9] = 0.f;
  acc[10] = -a.sum_a;
  acc[11] = -a.sum_b;
}

// scratch of tile (b, c): floats [kObsScratchPerTile]; the 12 float64 partial sums first (8-byte aligned), then (max, sum exp)
__device__ __forceinline__ float* obs_tile_scratch(float* scratch, int b, int c) {
  return scratch + ((long long)b * kObsChunks + c) * kObsScratchPerTile;
}

// row logsumexp from the tiles' (max, sum exp) pairs, in tile order
__device__ __forceinline__ float obs_row_lse(const float* scratch, int b) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kObsChunks; ++c) mx = fmaxf(mx, obs_tile_scratch(const_cast<float*>(scratch), b, c)[2 * kNumSums]);
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < kObsChunks; ++c) {
    const float* t = obs_tile_scratch(const_cast<float*>(scratch), b, c) + 2 * kNumSums;
    if (t[0] > -INFINITY) se += t[1] * exp_fast(t[0] - mx);
  }
  return mx + logf(se);
}

// ---- launch 1: (max, sum exp) of the logits of every (chunk, row) tile ----
__global__ void __launch_bounds__(kObsThreads) obs_row_lse_kernel(int G, int chunk_len, const float* __restrict__ logits,
                                                                   long long ldl, float* __restrict__ scratch) {
  __shared__ float red[32];
  const int c = blockIdx.x, b = blockIdx.y;
  const int g0 = c * chunk_len, g1 = min(G, g0 + chunk_len);
  const float* lrow = logits + (long long)b * ldl;
  const int n4 = g1 > g0 ? (g1 - g0) >> 2 : 0;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 l = *reinterpret_cast<const float4*>(lrow + g0 + 4 * i);
    mx = fmaxf(fmaxf(mx, fmaxf(l.x, l.y)), fmaxf(l.z, l.w));
  }
  for (int g = g0 + 4 * n4 + threadIdx.x; g < g1; g += blockDim.x) mx = fmaxf(mx, lrow[g]);
  mx = block_max(mx, red);
  float se[1] = {0.f};
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {  // second read of the tile: L1 / L2
    const float4 l = *reinterpret_cast<const float4*>(lrow + g0 + 4 * i);
    se[0] += exp_fast(l.x - mx) + exp_fast(l.y - mx) + exp_fast(l.z - mx) + exp_fast(l.w - mx);
  }
  for (int g = g0 + 4 * n4 + threadIdx.x; g < g1; g += blockDim.x) se[0] += exp_fast(lrow[g] - mx);
  block_sum<float, 1>(se, red);
  if (threadIdx.x == 0) {
    float* t = obs_tile_scratch(scratch, b, c) + 2 * kNumSums;
    t[0] = mx;  // -inf for an empty tile (g0 >= G)
    t[1] = se[0];
  }
}

// ---- launch 2: likelihood terms, their gradients and the partial row sums of every (chunk, row) tile ----
template <bool EXACT>
__global__ void __launch_bounds__(kObsThreads, CB_OBS_MIN_CTAS) obs_main_kernel(
    const long long* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
    const long long* __restrict__ row_ids, int G, int chunk_len, const float* __restrict__ logits, long long ldl,
    const float* __restrict__ chi_amb, const float* __restrict__ chi_bar, const float* __restrict__ coef,
    const float* __restrict__ alpha, const float* __restrict__ w, float* __restrict__ dlogits, float* __restrict__ qamb,
    long long ldq, float* __restrict__ scratch, int max_poisson) {
  extern __shared__ unsigned int bitmap[];  // chunk_len / 32 words: genes of the tile with a stored count
  __shared__ double dred[kNumSums * 32];
  __shared__ long long jrange[2];

  const int c_id = blockIdx.x, b = blockIdx.y;
  const int g0 = c_id * chunk_len, g1 = min(G, g0 + chunk_len);
  const long long r = row_ids[b];
  const long long s = indptr[r], e = indptr[r + 1];
  const float* lrow = logits + (long long)b * ldl;
  float* drow = dlogits + (long long)b * ldl;
  float* qrow = qamb + (long long)b * ldq;
  const int nwords = chunk_len >> 5;

  RowCoef c;
  c.a_mu1 = coef[b * 7 + 0], c.cA1 = coef[b * 7 + 1], c.cB1 = coef[b * 7 + 2], c.cA0 = coef[b * 7 + 3];
  c.cB0 = coef[b * 7 + 4], c.cAE = coef[b * 7 + 5], c.cBE = coef[b * 7 + 6];
  c.w1 = w[b * 3 + 0], c.w0 = w[b * 3 + 1], c.wE = w[b * 3 + 2];
  c.alpha = alpha[0];
  const float inv_alpha = 1.0f / c.alpha;
  const float k1 = c.w1 * c.cA1, q_lin = -(c.w0 * c.cA0 + c.wE * c.cAE);
  const float lse = obs_row_lse(scratch, b);

  // ---- stored counts of this tile: [js, je) of the row (sorted columns), bitmap of their genes ----
  if (threadIdx.x < 2) {
    const int key = threadIdx.x == 0 ? g0 : g1;
    long long lo = s, hi = e;  // first j with indices[j] >= key
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (indices[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    jrange[threadIdx.x] = lo;
  }
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0u;
  __syncthreads();
  const long long js = jrange[0], je = jrange[1];
  for (long long j = js + threadIdx.x; j < je; j += blockDim.x) {
    const int g = indices[j] - g0;
    atomicOr(&bitmap[g >> 5], 1u << (g & 31));
  }
  __syncthreads();

  // ---- dense pass: every gene of the tile without a stored count, x == 0 ----
  DenseAcc da = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int n4 = g1 > g0 ? (g1 - g0) >> 2 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int gl = 4 * i, gg = g0 + gl;
    const float4 l = *reinterpret_cast<const float4*>(lrow + gg);
    const float4 a4 = *reinterpret_cast<const float4*>(chi_amb + gg);
    const float4 b4 = *reinterpret_cast<const float4*>(chi_bar + gg);
    const unsigned int bits = (bitmap[gl >> 5] >> (gl & 31)) & 0xFu;
    const float ll[4] = {l.x, l.y, l.z, l.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float pp[4], qq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pp[k] = 0.f, qq[k] = 0.f;
      if (!((bits >> k) & 1u))
        obs_zero_element<EXACT>(exp_fast(ll[k] - lse), aa[k], bb[k], c, inv_alpha, k1, q_lin, da, pp[k], qq[k]);
    }
    *reinterpret_cast<float4*>(drow + gg) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(qrow + gg) = make_float4(qq[0], qq[1], qq[2], qq[3]);
  }
  for (int g = g0 + 4 * n4 + threadIdx.x; g < g1; g += blockDim.x) {
    float p = 0.f, q = 0.f;
    const int gl = g - g0;
    if (!((bitmap[gl >> 5] >> (gl & 31)) & 1u))
      obs_zero_element<EXACT>(exp_fast(lrow[g] - lse), chi_amb[g], chi_bar[g], c, inv_alpha, k1, q_lin, da, p, q);
    drow[g] = p;
    qrow[g] = q;
  }
  float acc[kNumSums];
  obs_dense_expand(da, c, acc);
  __syncthreads();  // dense stores to drow / qrow precede the sparse overwrites below

  // ---- sparse pass: the stored counts of the tile ----
  for (long long j = js + threadIdx.x; j < je; j += blockDim.x) {
    const int g = indices[j];
    float p, q;
    obs_element<EXACT>(data[j], expf(lrow[g] - lse), chi_amb[g], chi_bar[g], c, max_poisson, acc, p, q);
    drow[g] = p;
    qrow[g] = q;
  }

  // ---- partial row sums (double, fixed tree -> deterministic) ----
  double dacc[kNumSums];
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) dacc[i] = double(acc[i]);
  block_sum<double, kNumSums>(dacc, dred);
  if (threadIdx.x == 0) {
    // the + 1e-10 of lam_y0 and lam_E on the tile's zero-count genes (see obs_dense_expand)
    const double n_zero = double(g1 > g0 ? g1 - g0 : 0) - double(je - js);
    dacc[1] -= n_zero * (EXACT ? 2e-10 : 1e-10);
    dacc[2] -= n_zero * 1e-10;
    if (EXACT) dacc[9] -= n_zero * 0.5 * 1e-10 * 1e-10 * double(inv_alpha) * double(inv_alpha);
    double* out = reinterpret_cast<double*>(obs_tile_scratch(scratch, b, c_id));
#pragma unroll
    for (int i = 0; i < kNumSums; ++i) out[i] = dacc[i];
  }
}

// ---- launch 3: row sums from the tiles' partials; softmax backward  dlogit = w1 a_mu1 (p - chi * sum_g p) ----
__global__ void __launch_bounds__(kObsThreads) obs_softmax_bwd_kernel(int G, int chunk_len, const float* __restrict__ logits,
                                                                       long long ldl, const float* __restrict__ coef,
                                                                       const float* __restrict__ w, float* __restrict__ scratch,
                                                                       float* __restrict__ sums, float* __restrict__ dlogits,
                                                                       float* __restrict__ lse_out, int* __restrict__ nan_flag) {
  const int c_id = blockIdx.x, b = blockIdx.y;
  const int g0 = c_id * chunk_len, g1 = min(G, g0 + chunk_len);
  const float* lrow = logits + (long long)b * ldl;
  float* drow = dlogits + (long long)b * ldl;
  const float lse = obs_row_lse(scratch, b);
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < kObsChunks; ++k) tot += reinterpret_cast<const double*>(obs_tile_scratch(scratch, b, k))[3];
  const float s_mu_chi = float(tot);
  const float scale = w[b * 3 + 0] * coef[b * 7 + 0];  // w1 a_mu1
  if (c_id == 0 && threadIdx.x < kNumSums) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kObsChunks; ++k) t += reinterpret_cast<const double*>(obs_tile_scratch(scratch, b, k))[threadIdx.x];
    const float v = float(t);
    sums[b * kNumSums + threadIdx.x] = v;
    if (threadIdx.x < 3 && v != v) atomicOr(nan_flag, 1);
  }
  if (c_id == 0 && threadIdx.x == 0 && lse_out != nullptr) lse_out[b] = lse;
  const int n4 = g1 > g0 ? (g1 - g0) >> 2 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int gg = g0 + 4 * i;
    const float4 l = *reinterpret_cast<const float4*>(lrow + gg);
    float4 p = *reinterpret_cast<float4*>(drow + gg);
    p.x = scale * (p.x - exp_fast(l.x - lse) * s_mu_chi);
    p.y = scale * (p.y - exp_fast(l.y - lse) * s_mu_chi);
    p.z = scale * (p.z - exp_fast(l.z - lse) * s_mu_chi);
    p.w = scale * (p.w - exp_fast(l.w - lse) * s_mu_chi);
    *reinterpret_cast<float4*>(drow + gg) = p;
  }
  for (int g = g0 + 4 * n4 + threadIdx.x; g < g1; g += blockDim.x)
    drow[g] = scale * (drow[g] - exp_fast(lrow[g] - lse) * s_mu_chi);
}

// Deterministic column sum of a [R, ld] matrix: out[g] = sum_r in[r, g].  A CTA owns 32 adjacent columns at a time; its 256
// threads are 8 row groups x 32 columns (coalesced 128-byte row segments, 32 independent loads in flight per thread);
// the row groups are combined in shared memory in a fixed order.  The grid is capped at four CTAs per SM (column strips
// are strided over the CTAs): the one-CTA-per-strip launch filled every thread slot of every SM for its whole duration,
// and the small kernels of the step's critical chain that became ready beside it (the loss assembly behind the
// observation kernels, the decoder's input-gradient product) waited 10-15 us for a slot (r02j timeline).
constexpr int kColsumCols = 32, kColsumGroups = 8, kColsumUnroll = 32, kColsumCtasPerSm = 4;
__global__ void __launch_bounds__(kColsumCols* kColsumGroups) colsum_rows_kernel(const float* __restrict__ in, int R, int G,
                                                                                long long ld, float* __restrict__ out) {
  __shared__ float s[kColsumGroups][kColsumCols + 1];
  const int c = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int n_strips = (G + kColsumCols - 1) / kColsumCols;
  for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x) {
    const int g = strip * kColsumCols + c;
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
    __syncthreads();  // s is reused by the next strip
  }
}

}  // namespace cb

static int elbo_obs_launch(const long long* indptr, const int* indices, const float* data, const long long* row_ids, int n_rows,
                           int n_genes, const float* logits, long long ldl, const float* chi_ambient, const float* chi_bar,
                           const float* coef, const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                           long long ldq, float* lse_out, int* nan_flag, float* scratch, int max_poisson, void* stream) {
  if (n_rows == 0) return 0;
  CB_REQUIRE(n_rows > 0 && n_rows <= 65535 && n_genes > 0, "cb_elbo_obs_fwd_bwd: bad shape [%d, %d]", n_rows, n_genes);
  CB_REQUIRE((ldl & 3) == 0 && (ldq & 3) == 0 && ldl >= n_genes && ldq >= n_genes,
             "cb_elbo_obs_fwd_bwd: ldl=%lld / ldq=%lld must be >= n_genes and multiples of 4", ldl, ldq);
  for (const void* p : {(const void*)logits, (const void*)dlogits, (const void*)qamb, (const void*)chi_ambient,
                        (const void*)chi_bar})
    CB_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "cb_elbo_obs_fwd_bwd: pointer not 16-byte aligned");
  CB_REQUIRE(scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0,
             "cb_elbo_obs_fwd_bwd: scratch (cb_elbo_obs_scratch_elems floats, 8-byte aligned) is required");
  const int chunk_len = cb::obs_chunk_len(n_genes);
  const size_t smem = size_t(chunk_len / 32) * sizeof(unsigned int);
  CB_REQUIRE(smem <= 40 * 1024, "cb_elbo_obs_fwd_bwd: n_genes=%d too large for the shared-memory bitmap", n_genes);
  const dim3 grid(cb::kObsChunks, n_rows);
  cudaStream_t st = (cudaStream_t)stream;
  cb::obs_row_lse_kernel<<<grid, cb::kObsThreads, 0, st>>>(n_genes, chunk_len, logits, ldl, scratch);
  auto kern = max_poisson >= 0 ? cb::obs_main_kernel<true> : cb::obs_main_kernel<false>;
  kern<<<grid, cb::kObsThreads, smem, st>>>(indptr, indices, data, row_ids, n_genes, chunk_len, logits, ldl, chi_ambient, chi_bar,
                                            coef, alpha, w, dlogits, qamb, ldq, scratch, max_poisson);
  cb::obs_softmax_bwd_kernel<<<grid, cb::kObsThreads, 0, st>>>(n_genes, chunk_len, logits, ldl, coef, w, scratch, sums, dlogits,
                                                              lse_out, nan_flag);
  return cb::check_launch("cb_elbo_obs_fwd_bwd");
}

// floats of scratch the observation kernels need for n_rows droplets (partial row sums + logsumexp pairs per tile)
extern "C" long long cb_elbo_obs_scratch_elems(int n_rows) {
  return (long long)(n_rows > 0 ? n_rows : 0) * cb::kObsChunks * cb::kObsScratchPerTile;
}

extern "C" int cb_elbo_obs_fwd_bwd(const long long* indptr, const int* indices, const float* data,
                                   const long long* row_ids, int n_rows, int n_genes, const float* logits,
                                   long long ldl, const float* chi_ambient, const float* chi_bar, const float* coef,
                                   const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                                   long long ldq, float* lse_out, int* nan_flag, float* scratch, void* stream) {
  return elbo_obs_launch(indptr, indices, data, row_ids, n_rows, n_genes, logits, ldl, chi_ambient, chi_bar, coef, alpha, w, sums,
                         dlogits, qamb, ldq, lse_out, nan_flag, scratch, -1, stream);
}

// The same with the exact NB (+) Poisson convolution as the likelihood (NegativeBinomialPoissonConv, max_poisson terms).
extern "C" int cb_elbo_obs_exact_fwd_bwd(const long long* indptr, const int* indices, const float* data,
                                         const long long* row_ids, int n_rows, int n_genes, const float* logits,
                                         long long ldl, const float* chi_ambient, const float* chi_bar, const float* coef,
                                         const float* alpha, const float* w, float* sums, float* dlogits, float* qamb,
                                         long long ldq, float* lse_out, int* nan_flag, float* scratch, int max_poisson,
                                         void* stream) {
  CB_REQUIRE(max_poisson >= 0 && max_poisson <= 1000, "cb_elbo_obs_exact_fwd_bwd: max_poisson=%d not in [0, 1000]", max_poisson);
  return elbo_obs_launch(indptr, indices, data, row_ids, n_rows, n_genes, logits, ldl, chi_ambient, chi_bar, coef, alpha, w, sums,
                         dlogits, qamb, ldq, lse_out, nan_flag, scratch, max_poisson, stream);
}

extern "C" int cb_colsum_rows(const float* in, int n_rows, int n_cols, long long ld, float* out, void* stream) {
  CB_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "cb_colsum_rows: bad shape [%d, %d] ld=%lld", n_rows, n_cols, ld);
  const int n_strips = (n_cols + cb::kColsumCols - 1) / cb::kColsumCols, cap = cb::kColsumCtasPerSm * cb::sm_count();
  cb::colsum_rows_kernel<<<n_strips < cap ? n_strips : cap, cb::kColsumCols * cb::kColsumGroups, 0, (cudaStream_t)stream>>>(
      in, n_rows, n_cols, ld, out);
  return cb::check_launch("cb_colsum_rows");
}
