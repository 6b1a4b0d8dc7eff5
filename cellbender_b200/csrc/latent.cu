// Per-droplet latent algebra of one training step, forward and backward, and the simplex parameter transform.
//
// Replaces, per svi.step, everything in the reference's model()/guide() that lives on [B]- or [B, Z]-shaped tensors
// (reference cellbender/remove_background/model.py:232-445 model, :447-574 guide; vae/encoder.py:300-340 output
// shaping of EncodeNonZLatents) and the autograd graph pyro builds over it: ~900 tiny torch kernels per step in an
// eager restatement, here three launches of ONE CTA each (the work is B = 255 rows; it is latency, not throughput).
//
//   latent_prepare   encoder output shaping -> concentrations at which the Gamma / Beta base samples are drawn
//   latent_forward   constraint transforms of the scalar pyro.params, reparameterised samples
//                    (phi, rho, d_empty, p_logit_reg, z, d_cell, epsilon), rate coefficients for the fused
//                    observation kernel, enumeration weights, and every non-observation ELBO term
//   latent_backward  analytic gradient of -ELBO w.r.t. the encoder outputs and the scalar parameters, given the
//                    observation kernel's gradients (d coef, d alpha, d w) and the decoder's d z
//   simplex_forward / simplex_backward   transform_to(constraints.simplex) (torch SoftmaxTransform) of the
//                    chi_ambient parameter (model.py:245-249, 486-490) and its vector-Jacobian product
//
// The per-droplet algebra runs in fp32 registers (type R below), sums over droplets in fp64; inputs and outputs are fp32.
// The implicit reparameterisation gradients of the Gamma and Beta samples (torch._standard_gamma_grad,
// torch._dirichlet_grad -- what Gamma.rsample / Beta.rsample differentiate through) are passed in as factors.
#include "common.cuh"

namespace cb {

// Arithmetic type of the per-droplet algebra.  fp32 matches the reference (torch fp32) and keeps the one-CTA kernels
// short; sums over droplets are accumulated in fp64 either way.  -DCB_LATENT_FP64 switches everything to fp64.
#ifdef CB_LATENT_FP64
using R = double;
#else
using R = float;
#endif
constexpr R kHalfLog2Pi = R(0.91893853320467274178);
constexpr R kFltTiny = R(1.17549435e-38);
constexpr R kFltEps = R(1.1920928955078125e-07);

struct LatentConsts {
  float log_count_crossover, prior_log_cell_counts, empty_log_count_threshold, offset_logit_p;
  float d_cell_loc_prior, d_cell_scale_prior, epsilon_prior, phi_conc_prior, phi_rate_prior;
  float rho_alpha_prior, rho_beta_prior, d_empty_loc_prior, d_empty_scale_prior, p_logit_prior;
  float empty_umi_threshold, counts_crossover, rho_param_min, rho_param_max, model_type;
};
constexpr int kNumConsts = 19;

struct ScalarParams {  // unconstrained pyro.params (device pointers; rho_* may be null for model 'ambient')
  const float *d_cell_scale, *d_empty_loc, *d_empty_scale, *rho_alpha, *rho_beta, *phi_loc, *phi_scale;
};
struct ScalarGrads {
  float *d_cell_scale, *d_empty_loc, *d_empty_scale, *rho_alpha, *rho_beta, *phi_loc, *phi_scale;
};

__device__ __forceinline__ R sigmoid_d(R x) { return R(1.0) / (R(1.0) + exp(-x)); }
__device__ __forceinline__ R logsigmoid_d(R x) { return fmin(x, R(0.0)) - log1p(exp(-fabs(x))); }
// torch.nn.Softplus(beta=1, threshold=20) and its derivative
__device__ __forceinline__ R softplus_d(R x) { return x > R(20.0) ? x : log1p(exp(x)); }
__device__ __forceinline__ R softplus_grad_d(R x) { return x > R(20.0) ? R(1.0) : sigmoid_d(x); }
__device__ __forceinline__ R clipped_sigmoid_d(R x) {
  return fmin(fmax(sigmoid_d(x), kFltTiny), R(1.0) - kFltEps);  // torch.distributions.transforms._clipped_sigmoid (fp32)
}
__device__ R digamma_d(R x) {
  R r = R(0.0);
  while (x < R(6.0)) {
    r -= R(1.0) / x;
    x += R(1.0);
  }
  const R f = R(1.0) / (x * x);
  return r + log(x) - R(0.5) / x - f * (R(1.0) / R(12.0) - f * (R(1.0) / R(120.0) - f * (R(1.0) / R(252.0) - f * (R(1.0) / R(240.0) - f / R(132.0)))));
}
__device__ __forceinline__ R xlogy_d(R x, R y) { return x == R(0.0) ? R(0.0) : x * log(y); }

struct Globals {  // constrained scalar parameters and what derives from them
  R dcs, del, des, ra, rb, pl, psc, phi_conc, phi_rate;
  R sra, srb;  // clipped sigmoids behind rho_alpha / rho_beta
  R d_empty_enc;  // exp(d_empty_loc): what EncodeNonZLatents subtracts from the counts (vae/encoder.py:330)
  bool has_rho;
};

__device__ Globals load_globals(const ScalarParams& p, const LatentConsts& k) {
  Globals g;
  g.dcs = exp(R(p.d_cell_scale[0]));
  g.del = exp(R(p.d_empty_loc[0]));
  g.des = exp(R(p.d_empty_scale[0]));
  g.has_rho = (p.rho_alpha != nullptr) && (k.model_type != 0.f);
  g.sra = g.srb = R(0.0);
  g.ra = g.rb = R(1.0);
  if (g.has_rho) {
    const R lo = k.rho_param_min, hi = k.rho_param_max;
    g.sra = clipped_sigmoid_d(R(p.rho_alpha[0]));
    g.srb = clipped_sigmoid_d(R(p.rho_beta[0]));
    g.ra = lo + (hi - lo) * g.sra;
    g.rb = lo + (hi - lo) * g.srb;
  }
  g.d_empty_enc = exp(g.del);
  g.pl = exp(R(p.phi_loc[0]));
  g.psc = exp(R(p.phi_scale[0]));
  g.phi_conc = (g.pl / g.psc) * (g.pl / g.psc);
  g.phi_rate = g.pl / (g.psc * g.psc);
  return g;
}

struct EncRow {  // EncodeNonZLatents output shaping (vae/encoder.py:300-340) for one droplet
  R counts, p_y, dpy_dpout, eps_enc, deps_dout, d_loc, ddloc_deps, prob, q1, q0, lq1, lq0;
};

__device__ EncRow shape_encoder_row(float counts_f, float p_out, float eps_out, R d_empty_enc, const LatentConsts& k) {
  EncRow r;
  const R c = counts_f;
  r.counts = c;
  const R ls = log1p(c);
  const R dlt = ls - R(k.log_count_crossover);
  const R sgn = (dlt > R(0.0)) - (dlt < R(0.0));
  R p = (R(p_out) - R(k.offset_logit_p)) + sqrt(fabs(dlt)) * sgn * R(10.0);
  const R a1 = sigmoid_d(R(50.0) * (ls - R(k.prior_log_cell_counts)));
  p = (R(1.0) - a1) * p + a1 * R(10.0);
  const R a0 = sigmoid_d(R(50.0) * (ls - R(k.empty_log_count_threshold)));
  r.p_y = a0 * p + (R(1.0) - a0) * (-R(10.0));
  r.dpy_dpout = a0 * (R(1.0) - a1);
  const R s = sigmoid_d(R(eps_out) * R(0.2) - R(1.0986122886681098));
  r.eps_enc = R(2.0) * s + R(0.5);
  r.deps_dout = R(2.0) * s * (R(1.0) - s) * R(0.2);
  const R u = c / (r.eps_enc + R(1e-2)) - d_empty_enc;
  const R spu = softplus_d(u) + R(1e-10);
  const R t = log(spu) - R(k.log_count_crossover);
  r.d_loc = softplus_d(t) + R(k.log_count_crossover);
  r.ddloc_deps = softplus_grad_d(t) * (R(1.0) / spu) * softplus_grad_d(u) * (-c / ((r.eps_enc + R(1e-2)) * (r.eps_enc + R(1e-2))));
  r.prob = sigmoid_d(r.p_y);
  r.lq1 = logsigmoid_d(r.p_y);
  r.lq0 = logsigmoid_d(-r.p_y);
  r.q1 = exp(r.lq1);
  r.q0 = exp(r.lq0);
  return r;
}

__device__ __forceinline__ R p_logit_prior_row(R counts, const LatentConsts& k) {  // model.py:577-592
  const R lc = log(counts);
  const R thresh = (log(R(k.empty_umi_threshold)) + R(k.d_cell_loc_prior)) / R(2.0);
  R p = k.p_logit_prior;
  if (lc <= thresh) p = -R(100.0);
  if (lc >= R(k.d_cell_loc_prior)) p = R(10.0);
  return p;
}

struct SampleRow {  // the reparameterised guide samples of one droplet
  R rho, lde, d_empty, lx, d_cell, eps_conc, epsilon, n_dc, n_de, n_v;
};

__device__ SampleRow sample_row(const EncRow& e, const Globals& g, const LatentConsts& k, const float* n_row3, float g_eps,
                                const float* xx2) {
  SampleRow s;
  s.n_dc = n_row3[0], s.n_de = n_row3[1], s.n_v = n_row3[2];
  s.rho = g.has_rho ? R(xx2[0]) : R(0.0);
  s.lde = g.del + g.des * s.n_de;
  s.d_empty = exp(s.lde);
  const R gated = e.prob * e.d_loc + (R(1.0) - e.prob) * R(k.d_cell_loc_prior);
  s.lx = gated + g.dcs * s.n_dc;
  s.d_cell = exp(s.lx);
  s.eps_conc = (e.prob * e.eps_enc + (R(1.0) - e.prob)) * R(k.epsilon_prior);
  s.epsilon = fmax(R(g_eps) / R(k.epsilon_prior), kFltTiny);
  return s;
}

// block-wide sum of N Rs per thread (result in every thread); smem: N * 32 Rs
template <int N>
__device__ void block_sum_d(double (&v)[N], double* smem) { block_sum<double, N>(v, smem); }

// Lower medians (torch.median) of vals over the two disjoint droplet classes mask bit 1 (probable empties) and bit 2
// (probable cells): index of the element of rank (k-1)/2 in (value, index) order within its class, -1 for an empty
// class.  Every droplet belongs to exactly one class.  Each droplet gets ONE sortable 64-bit key (class, value bits in
// total order, index), so its rank inside its class is (number of smaller keys) - (size of the lower class): the
// rank-counting sweep is one 64-bit compare per pair, split over all threads of the CTA (B x P work items, P = threads
// / B pieces of the j range per droplet).  r02x profile: the former per-pair class / value / index tests were 30 % of the
// forward kernel's instructions on half of its warps.
// res[0] = cells, res[1] = empties, res[2] = number of droplets with mask bit 4 (surely cells).
__device__ void class_medians(const float* vals, const unsigned char* mask, int B, int* res, int* cnt,
                              unsigned long long* keys, int* rank_s) {
  if (threadIdx.x < 3) res[threadIdx.x] = (threadIdx.x == 2) ? 0 : -1;
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  int c_cell = 0, c_empty = 0, c_sure = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const unsigned char m = mask[i];
    c_cell += (m & 2) ? 1 : 0, c_empty += (m & 1) ? 1 : 0, c_sure += (m & 4) ? 1 : 0;
    unsigned int bits = __float_as_uint(vals[i]);
    bits = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);  // unsigned order == float order
    const unsigned long long cls = ((m & 3) == 2) ? 0ull : 1ull;  // cells sort below empties
    keys[i] = (cls << 63) | ((unsigned long long)bits << 16) | (unsigned long long)i;
    rank_s[i] = 0;
  }
  if (c_cell) atomicAdd(&cnt[0], c_cell);
  if (c_empty) atomicAdd(&cnt[1], c_empty);
  if (c_sure) atomicAdd(&res[2], c_sure);
  __syncthreads();
  const int P = (int(blockDim.x) / B) > 1 ? int(blockDim.x) / B : 1;
  const int len = (B + P - 1) / P;
  for (int w = threadIdx.x; w < B * P; w += blockDim.x) {
    const int i = w % B, part = w / B;
    const unsigned long long ki = keys[i];
    const int j1 = (part + 1) * len < B ? (part + 1) * len : B;
    int c = 0;
#pragma unroll 4
    for (int j = part * len; j < j1; ++j) c += keys[j] < ki;
    if (c) atomicAdd(&rank_s[i], c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int which = int(keys[i] >> 63);
    const int rank = rank_s[i] - (which ? cnt[0] : 0);
    if (rank == (cnt[which] - 1) / 2) res[which] = i;
  }
  __syncthreads();
}

// ---- z = loc + exp(log_scale) * n over [B, Z]: work distribution of the vector path --------------------------------
// Z % 4 == 0 and every row 16-byte aligned: LPR lanes (a power of two, Z / 4 rounded up, at most 32) share one droplet,
// each lane walks float4 columns; a warp covers 32 / LPR droplets at once and the row sums need log2(LPR) shuffles.
// (Z = 64: 16 lanes per droplet, one float4 each -- a quarter of the instructions of the 32-wide scalar slices.)
struct ZLanes {
  int lpr, sub, rw, rows_per_warp;
};
__device__ __forceinline__ ZLanes z_lanes(int Z) {
  ZLanes t;
  const int z4 = Z >> 2;
  t.lpr = 1;
  while (t.lpr < z4 && t.lpr < 32) t.lpr <<= 1;
  const int lane = threadIdx.x & 31;
  t.sub = lane & (t.lpr - 1);
  t.rw = lane / t.lpr;
  t.rows_per_warp = 32 / t.lpr;
  return t;
}
__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_prepare_kernel(const float* __restrict__ feats, const float* __restrict__ p_out,
                                                             const float* __restrict__ eps_out, int B, ScalarParams sp,
                                                             LatentConsts k, float* __restrict__ conc_gamma,
                                                             float* __restrict__ conc_beta, float* __restrict__ total_beta) {
  const Globals g = load_globals(sp, k);
  if (threadIdx.x == 0 && blockIdx.x == 0) conc_gamma[0] = float(g.phi_conc);
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const EncRow e = shape_encoder_row(feats[4 * b], p_out[b], eps_out[b], g.d_empty_enc, k);
    conc_gamma[1 + b] = float((e.prob * e.eps_enc + (R(1.0) - e.prob)) * R(k.epsilon_prior));
    conc_gamma[1 + B + b] = float(g.ra);
    conc_gamma[1 + 2 * B + b] = float(g.rb);
    conc_beta[2 * b] = float(g.ra), conc_beta[2 * b + 1] = float(g.rb);
    total_beta[2 * b] = total_beta[2 * b + 1] = float(g.ra + g.rb);
  }
}

// standard-gamma draws -> clamp to >= tiny (Gamma.rsample) and rho = ga / (ga + gb) as the Dirichlet pair (x, 1 - x)
__global__ void latent_beta_kernel(float* __restrict__ gdraw, int B, float* __restrict__ xx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1 + 3 * B) gdraw[i] = fmaxf(gdraw[i], 1.17549435e-38f);
  if (i < B) {
    const R ga = fmax(R(gdraw[1 + B + i]), kFltTiny), gb = fmax(R(gdraw[1 + 2 * B + i]), kFltTiny);
    R x = ga / (ga + gb);
    x = fmin(fmax(x, kFltTiny), R(1.0) - kFltEps);
    xx[2 * i] = float(x);
    xx[2 * i + 1] = float(R(1.0) - x);
  }
}

// dynamic shared memory: float zpart_p[B * nchunk], zpart_q[B * nchunk], eps[B], q1[B]; uchar mask[B]; (8-byte aligned)
// u64 keys[B]; int rank[B]
__host__ __device__ inline size_t latent_keys_offset(int B, int Z) {
  return (size_t(B) * ((2 * ((Z + 31) / 32) + 2) * sizeof(float) + 1) + 15) / 16 * 16;
}

enum Term { T_PHI = 0, T_Z, T_DCELL, T_RHO, T_EPS, T_DEMPTY, T_Y, T_PLR, T_SOFT_EMPTY, T_SOFT_CELL, T_EPS_MEAN, T_EPS_EMPTY_MEAN, T_N };

constexpr int kLatentThreads = 512, kZUnroll = 8;
__global__ void __launch_bounds__(kLatentThreads) latent_forward_kernel(
    const float* __restrict__ feats, const float* __restrict__ p_out, const float* __restrict__ eps_out,
    const float* __restrict__ z_loc, const float* __restrict__ z_ls, long long ldz, int B, int Z,
    const float* __restrict__ n_z, const float* __restrict__ n_row, const float* __restrict__ g_gamma,
    const float* __restrict__ xx, ScalarParams sp, LatentConsts k, float* __restrict__ z_out, float* __restrict__ coef,
    float* __restrict__ alpha_out, float* __restrict__ w_out, float* __restrict__ terms, float* __restrict__ loss_local,
    float* __restrict__ enc_out, int* __restrict__ med_idx, int vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nchunk = (Z + 31) >> 5;
  float* zpart_p = reinterpret_cast<float*>(smem_raw);  // [B * nchunk] partial sums of the prior log-density
  float* zpart_q = zpart_p + B * nchunk;                // [B * nchunk] partial sums of the guide log-density
  float* eps_s = zpart_q + B * nchunk;
  unsigned char* mask = reinterpret_cast<unsigned char*>(eps_s + 2 * B);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + latent_keys_offset(B, Z));
  int* rank_s = reinterpret_cast<int*>(keys + B);
  __shared__ double red[T_N * 32];
  __shared__ int s_res[3], s_cnt[2];

  const Globals g = load_globals(sp, k);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;

  // ---- z = loc + exp(log_scale) * n; sums of the prior and guide log-densities (model.py:262-264, 553-555).
  const int zn = vec ? 1 : nchunk;  // partial sums per droplet
  if (vec) {
    const ZLanes t = z_lanes(Z);
    const int z4 = Z >> 2;
    constexpr int U = 2;  // droplet groups per iteration: their loads are issued before the first use
    for (int r0 = wid * t.rows_per_warp; r0 < B; r0 += nw * t.rows_per_warp * U) {
      float sp_[U], sq_[U];
#pragma unroll
      for (int i = 0; i < U; ++i) sp_[i] = sq_[i] = 0.f;
      for (int c = t.sub; c < z4; c += t.lpr) {
        float4 nn[U], ll[U], mm[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int r = r0 + i * nw * t.rows_per_warp + t.rw;
          if (r < B) {
            nn[i] = *reinterpret_cast<const float4*>(n_z + (long long)r * Z + 4 * c);
            ll[i] = *reinterpret_cast<const float4*>(z_ls + r * ldz + 4 * c);
            mm[i] = *reinterpret_cast<const float4*>(z_loc + r * ldz + 4 * c);
          }
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int r = r0 + i * nw * t.rows_per_warp + t.rw;
          if (r < B) {
            const float* n4 = reinterpret_cast<const float*>(&nn[i]);
            const float* l4 = reinterpret_cast<const float*>(&ll[i]);
            const float* m4 = reinterpret_cast<const float*>(&mm[i]);
            float4 zo;
            float* z4o = reinterpret_cast<float*>(&zo);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const R z = R(m4[e]) + exp(R(l4[e])) * R(n4[e]);
              z4o[e] = float(z);
              sp_[i] += float(-R(0.5) * z * z - kHalfLog2Pi);
              sq_[i] += float(-R(0.5) * R(n4[e]) * R(n4[e]) - R(l4[e]) - kHalfLog2Pi);
            }
            *reinterpret_cast<float4*>(z_out + (long long)r * Z + 4 * c) = zo;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int r = r0 + i * nw * t.rows_per_warp + t.rw;
        const float a = group_sum(sp_[i], t.lpr), b2 = group_sum(sq_[i], t.lpr);
        if (r < B && t.sub == 0) zpart_p[r] = a, zpart_q[r] = b2;
      }
    }
  } else {
    // scalar path (Z % 4 != 0 or unaligned rows): work unit = (droplet, 32-wide slice of z); every warp keeps kZUnroll
    // units' loads in flight
    const int n_units = B * nchunk;
    for (int u0 = wid * kZUnroll; u0 < n_units; u0 += nw * kZUnroll) {
      float nn[kZUnroll], ll[kZUnroll], mm[kZUnroll];
#pragma unroll
      for (int i = 0; i < kZUnroll; ++i) {
        const int u = u0 + i, r = u / nchunk, j = (u - r * nchunk) * 32 + lane;
        const bool in = u < n_units && j < Z;
        nn[i] = in ? n_z[(long long)r * Z + j] : 0.f;
        ll[i] = in ? z_ls[r * ldz + j] : 0.f;
        mm[i] = in ? z_loc[r * ldz + j] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kZUnroll; ++i) {
        const int u = u0 + i, r = u / nchunk, j = (u - r * nchunk) * 32 + lane;
        const bool in = u < n_units && j < Z;
        const R z = R(mm[i]) + exp(R(ll[i])) * R(nn[i]);
        if (in) z_out[(long long)r * Z + j] = float(z);
        float sp_ = in ? float(-R(0.5) * z * z - kHalfLog2Pi) : 0.f;
        float sq_ = in ? float(-R(0.5) * R(nn[i]) * R(nn[i]) - R(ll[i]) - kHalfLog2Pi) : 0.f;
        sp_ = warp_sum(sp_), sq_ = warp_sum(sq_);
        if (lane == 0 && u < n_units) zpart_p[u] = sp_, zpart_q[u] = sq_;
      }
    }
  }
  __syncthreads();

  double acc[T_N];
#pragma unroll
  for (int i = 0; i < T_N; ++i) acc[i] = R(0.0);
  const R ep = k.epsilon_prior;
  const R phi = fmax(R(g_gamma[0]) / g.phi_rate, kFltTiny);

  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const EncRow e = shape_encoder_row(feats[4 * b], p_out[b], eps_out[b], g.d_empty_enc, k);
    const SampleRow s = sample_row(e, g, k, n_row + 3 * b, g_gamma[1 + b], xx + 2 * b);
    enc_out[4 * b + 0] = float(e.p_y), enc_out[4 * b + 1] = float(e.d_loc), enc_out[4 * b + 2] = float(e.eps_enc);
    enc_out[4 * b + 3] = float(s.epsilon);
    // rate coefficients (calculate_mu / calculate_lambda, model.py:25-85, factored; obs_empty: rho, d_cell detached, eps = 1)
    const R r = s.rho, ee = s.epsilon, dc = s.d_cell, de = s.d_empty;
    const R amb = (k.model_type == 1.f) ? R(0.0) : R(1.0);  // 'swapping' has no chi_ambient term
    float* cf = coef + 7 * b;
    cf[0] = float((R(1.0) - r) * ee * dc);
    cf[1] = float(amb * ee * (R(1.0) - r) * de);
    cf[2] = float(ee * r * (dc + de));
    cf[3] = float(amb * ee * (R(1.0) - r) * de);
    cf[4] = float(ee * r * de);
    cf[5] = float(amb * (R(1.0) - r) * de);
    cf[6] = float(r * de);
    w_out[3 * b + 0] = float(-e.q1), w_out[3 * b + 1] = float(-e.q0), w_out[3 * b + 2] = float(-R(0.01) * e.q0);

    {
      R zp = R(0.0), zq = R(0.0);
      for (int c2 = 0; c2 < zn; ++c2) zp += R(zpart_p[b * zn + c2]), zq += R(zpart_q[b * zn + c2]);
      acc[T_Z] += zp - e.q1 * zq;
    }
    {  // d_cell: LogNormal prior vs guide (the -log d_cell of both LogNormal densities cancels)
      const R dz = (s.lx - R(k.d_cell_loc_prior)) / R(k.d_cell_scale_prior);
      acc[T_DCELL] += (-R(0.5) * dz * dz - log(R(k.d_cell_scale_prior))) - (-R(0.5) * s.n_dc * s.n_dc - log(g.dcs));
    }
    if (g.has_rho) {
      const R a0 = k.rho_alpha_prior, b0 = k.rho_beta_prior, lr = log(r), l1r = log1p(-r);
      // the log-normalisers of both Beta densities are the same for every droplet: added once below (7 lgammas per
      // droplet were a fifth of this kernel's instructions)
      const R prior = xlogy_d(a0 - R(1.0), r) + (b0 - R(1.0)) * l1r;
      const R guide = (g.ra - R(1.0)) * lr + (g.rb - R(1.0)) * l1r;
      acc[T_RHO] += prior - guide;
    }
    {
      const R le = log(ee);
      const R prior = ep * log(ep) + (ep - R(1.0)) * le - ep * ee;  // - lgamma(ep): added once below
      const R guide = s.eps_conc * log(ep) + (s.eps_conc - R(1.0)) * le - ep * ee - lgamma(s.eps_conc);
      acc[T_EPS] += prior - guide;
    }
    {
      const R dz = (s.lde - R(k.d_empty_loc_prior)) / R(k.d_empty_scale_prior);
      acc[T_DEMPTY] += (-R(0.5) * dz * dz - log(R(k.d_empty_scale_prior))) - (-R(0.5) * s.n_de * s.n_de - log(g.des));
    }
    {
      const R plp = p_logit_prior_row(e.counts, k);
      acc[T_Y] += e.q1 * (logsigmoid_d(plp) - e.lq1) + e.q0 * (logsigmoid_d(-plp) - e.lq0);
      const R v = e.p_y + R(2.0) * s.n_v, dv = (v - R(k.p_logit_prior)) / R(2.0);
      acc[T_PLR] += (-R(0.5) * dv * dv) - (-R(0.5) * s.n_v * s.n_v);
    }
    const bool is_empty = e.counts < R(k.counts_crossover);
    {  // soft semi-supervision of p_y (model.py:400-427), Normal(+-10, 1).log_prob(p_y), scale 10
      const R d = is_empty ? (e.p_y + R(10.0)) : (e.p_y - R(10.0));
      acc[is_empty ? T_SOFT_EMPTY : T_SOFT_CELL] += R(10.0) * (-R(0.5) * d * d - kHalfLog2Pi);
    }
    eps_s[b] = float(ee);
    const bool surely_cell = e.counts >= exp(R(k.d_cell_loc_prior));
    mask[b] = (unsigned char)((is_empty ? 1 : 2) | (surely_cell ? 4 : 0));
  }
  {  // droplet-independent log-normalisers, B times each, one lgamma per thread of the last warp (idle in the row loop
     // when B < threads)
    const int t = int(blockDim.x) - 1 - int(threadIdx.x);
    if (t < 7) {
      const R a0 = k.rho_alpha_prior, b0 = k.rho_beta_prior;
      const R arg = t == 0 ? a0 + b0 : t == 1 ? a0 : t == 2 ? b0 : t == 3 ? g.ra + g.rb : t == 4 ? g.ra : t == 5 ? g.rb : ep;
      const double sign = (t == 0 || t == 4 || t == 5) ? 1.0 : -1.0;  // prior - guide
      if (t == 6) acc[T_EPS] += -double(B) * double(lgamma(arg));
      else if (g.has_rho) acc[T_RHO] += sign * double(B) * double(lgamma(arg));
    }
  }
  if (threadIdx.x == 0) {
    const R pc0 = k.phi_conc_prior, pr0 = k.phi_rate_prior, lphi = log(phi);
    const R prior = pc0 * log(pr0) + (pc0 - R(1.0)) * lphi - pr0 * phi - lgamma(pc0);
    const R guide = g.phi_conc * log(g.phi_rate) + (g.phi_conc - R(1.0)) * lphi - g.phi_rate * phi - lgamma(g.phi_conc);
    acc[T_PHI] = prior - guide;
    alpha_out[0] = float(R(1.0) / phi + R(1e-10));
  }
  __syncthreads();
  // ---- epsilon medians over probable cells / probable empties (model.py:429-443)
  class_medians(eps_s, mask, B, s_res, s_cnt, keys, rank_s);
  if (threadIdx.x == 0) {
    const int i_cell = s_res[0], i_empty = s_res[1], n_surely = s_res[2];
    const R lognorm = -log(R(0.01)) - kHalfLog2Pi;
    const R mc = (i_cell >= 0) ? R(eps_s[i_cell]) : R(1.0), me = (i_empty >= 0) ? R(eps_s[i_empty]) : R(1.0);
    const bool use_cell = n_surely >= 2;
    acc[T_EPS_MEAN] = use_cell ? (R(-0.5) * (R(1.0) - mc) * (R(1.0) - mc) / R(1e-4) + lognorm) : R(0.0);
    acc[T_EPS_EMPTY_MEAN] = R(-0.5) * (R(1.0) - me) * (R(1.0) - me) / R(1e-4) + lognorm;
    med_idx[0] = (use_cell ? i_cell : -1);
    med_idx[1] = i_empty;
  }
  block_sum_d<T_N>(acc, red);
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < T_N; ++i) {
      terms[i] = float(acc[i]);
      tot += acc[i];
    }
    loss_local[0] = float(-tot);  // contribution of these terms to the loss (-ELBO)
  }
}

__global__ void __launch_bounds__(kLatentThreads) latent_backward_kernel(
    const float* __restrict__ feats, const float* __restrict__ p_out, const float* __restrict__ eps_out,
    const float* __restrict__ z_loc, const float* __restrict__ z_ls, long long ldz, int B, int Z,
    const float* __restrict__ n_z, const float* __restrict__ n_row, const float* __restrict__ g_gamma,
    const float* __restrict__ xx, const float* __restrict__ sgg, const float* __restrict__ dgrad,
    const int* __restrict__ med_idx, ScalarParams sp, LatentConsts k, const float* __restrict__ d_z,
    const float* __restrict__ d_coef, const float* __restrict__ d_alpha, const float* __restrict__ d_w,
    const float* __restrict__ g_local, float* __restrict__ d_p_out, float* __restrict__ d_eps_out,
    float* __restrict__ d_z_loc, float* __restrict__ d_z_ls, long long lddz, ScalarGrads sg, int vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nchunk = (Z + 31) >> 5;
  float* zpart_q = reinterpret_cast<float*>(smem_raw) + B * nchunk;  // same layout as the forward kernel
  float* q1_s = zpart_q + B * nchunk + B;
  __shared__ double red[7 * 32];

  const Globals g = load_globals(sp, k);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const R gl = g_local ? R(g_local[0]) : R(1.0);  // upstream gradient of the local-terms loss output

  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const EncRow e = shape_encoder_row(feats[4 * b], p_out[b], eps_out[b], g.d_empty_enc, k);
    q1_s[b] = float(e.q1);
  }
  __syncthreads();
  // ---- z: d loss / d loc, d loss / d log_scale; guide log-density sums for the d q1 path (work split as in the forward)
  const int zn = vec ? 1 : nchunk;
  if (vec) {
    const ZLanes t = z_lanes(Z);
    const int z4 = Z >> 2;
    constexpr int U = 2;
    for (int r0 = wid * t.rows_per_warp; r0 < B; r0 += nw * t.rows_per_warp * U) {
      float sq_[U];
#pragma unroll
      for (int i = 0; i < U; ++i) sq_[i] = 0.f;
      for (int c = t.sub; c < z4; c += t.lpr) {
        float4 nn[U], ll[U], mm[U], dd[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int r = r0 + i * nw * t.rows_per_warp + t.rw;
          dd[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < B) {
            nn[i] = *reinterpret_cast<const float4*>(n_z + (long long)r * Z + 4 * c);
            ll[i] = *reinterpret_cast<const float4*>(z_ls + r * ldz + 4 * c);
            mm[i] = *reinterpret_cast<const float4*>(z_loc + r * ldz + 4 * c);
            if (d_z) dd[i] = *reinterpret_cast<const float4*>(d_z + (long long)r * Z + 4 * c);
          }
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int r = r0 + i * nw * t.rows_per_warp + t.rw;
          if (r < B) {
            const float* n4 = reinterpret_cast<const float*>(&nn[i]);
            const float* l4 = reinterpret_cast<const float*>(&ll[i]);
            const float* m4 = reinterpret_cast<const float*>(&mm[i]);
            const float* d4 = reinterpret_cast<const float*>(&dd[i]);
            const R q1 = R(q1_s[r]);
            float4 o_loc, o_ls;
            float* ol = reinterpret_cast<float*>(&o_loc);
            float* os = reinterpret_cast<float*>(&o_ls);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const R n = n4[e], ls = l4[e], sc = exp(ls);
              const R z = R(m4[e]) + sc * n;
              const R gz = gl * z + R(d4[e]);  // -d/dz of the N(0,1) prior term + observation path
              ol[e] = float(gz);
              os[e] = float(gz * n * sc - gl * q1);
              sq_[i] += float(-R(0.5) * n * n - ls - kHalfLog2Pi);
            }
            *reinterpret_cast<float4*>(d_z_loc + r * lddz + 4 * c) = o_loc;
            *reinterpret_cast<float4*>(d_z_ls + r * lddz + 4 * c) = o_ls;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int r = r0 + i * nw * t.rows_per_warp + t.rw;
        const float a = group_sum(sq_[i], t.lpr);
        if (r < B && t.sub == 0) zpart_q[r] = a;
      }
    }
  } else {
    const int n_units = B * nchunk;
    for (int u0 = wid * kZUnroll; u0 < n_units; u0 += nw * kZUnroll) {
      float nn[kZUnroll], ll[kZUnroll], mm[kZUnroll], dd[kZUnroll];
#pragma unroll
      for (int i = 0; i < kZUnroll; ++i) {
        const int u = u0 + i, r = u / nchunk, j = (u - r * nchunk) * 32 + lane;
        const bool in = u < n_units && j < Z;
        nn[i] = in ? n_z[(long long)r * Z + j] : 0.f;
        ll[i] = in ? z_ls[r * ldz + j] : 0.f;
        mm[i] = in ? z_loc[r * ldz + j] : 0.f;
        dd[i] = (in && d_z) ? d_z[(long long)r * Z + j] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kZUnroll; ++i) {
        const int u = u0 + i, r = u / nchunk, j = (u - r * nchunk) * 32 + lane;
        const bool in = u < n_units && j < Z;
        const R n = nn[i], ls = ll[i], sc = exp(ls);
        const R z = R(mm[i]) + sc * n;
        const R gz = gl * z + R(dd[i]);  // -d/dz of the N(0,1) prior term + observation path
        if (in) {
          d_z_loc[r * lddz + j] = float(gz);
          d_z_ls[r * lddz + j] = float(gz * n * sc - gl * R(q1_s[r]));
        }
        float sq_ = in ? float(-R(0.5) * n * n - ls - kHalfLog2Pi) : 0.f;
        sq_ = warp_sum(sq_);
        if (lane == 0 && u < n_units) zpart_q[u] = sq_;
      }
    }
  }
  __syncthreads();

  double ga[7];  // d loss / d (d_cell_scale, d_empty_loc, d_empty_scale, rho_alpha, rho_beta, phi_loc, phi_scale), constrained
#pragma unroll
  for (int i = 0; i < 7; ++i) ga[i] = R(0.0);
  const R ep = k.epsilon_prior;
  const int i_cell = med_idx[0], i_empty = med_idx[1];

  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const EncRow e = shape_encoder_row(feats[4 * b], p_out[b], eps_out[b], g.d_empty_enc, k);
    const SampleRow s = sample_row(e, g, k, n_row + 3 * b, g_gamma[1 + b], xx + 2 * b);
    const R r = s.rho, ee = s.epsilon, dc = s.d_cell, de = s.d_empty;
    const bool swapping = (k.model_type == 1.f);
    R G0 = 0, G1 = 0, G2 = 0, G3 = 0, G4 = 0, G5 = 0, G6 = 0;
    if (d_coef) {
      const float* gc = d_coef + 7 * b;
      G0 = gc[0], G2 = gc[2], G4 = gc[4], G6 = gc[6];
      if (!swapping) G1 = gc[1], G3 = gc[3], G5 = gc[5];
    }
    const R G13 = G1 + G3;
    R G_e = G0 * (R(1.0) - r) * dc + G13 * (R(1.0) - r) * de + G2 * r * (dc + de) + G4 * r * de;
    R G_r = -G0 * ee * dc - G13 * ee * de + G2 * ee * (dc + de) + G4 * ee * de;
    const R G_dc = G0 * (R(1.0) - r) * ee + G2 * ee * r;
    const R G_de = G13 * ee * (R(1.0) - r) + (G2 + G4) * ee * r + G5 * (R(1.0) - r) + G6 * r;
    // d_cell
    R G_lx = G_dc * dc + gl * (s.lx - R(k.d_cell_loc_prior)) / (R(k.d_cell_scale_prior) * R(k.d_cell_scale_prior));
    ga[0] += G_lx * s.n_dc - gl / g.dcs;
    const R G_dloc = G_lx * e.prob;
    // d_empty
    const R G_lde = G_de * de + gl * (s.lde - R(k.d_empty_loc_prior)) / (R(k.d_empty_scale_prior) * R(k.d_empty_scale_prior));
    ga[1] += G_lde;
    ga[2] += G_lde * s.n_de - gl / g.des;
    // rho
    if (g.has_rho) {
      const R a0 = k.rho_alpha_prior, b0 = k.rho_beta_prior;
      const R dt_dr = (a0 - R(1.0)) / r - (b0 - R(1.0)) / (R(1.0) - r) - (g.ra - R(1.0)) / r + (g.rb - R(1.0)) / (R(1.0) - r);
      G_r += -gl * dt_dr;
      // d/d(alpha, beta) of the guide's log-normaliser, psi(a + b) - psi(a | b), is the same for every droplet: added
      // once below
      ga[3] += gl * log(r) + G_r * R(dgrad[2 * b]) * (R(1.0) - r);
      ga[4] += gl * log1p(-r) - G_r * R(dgrad[2 * b + 1]) * r;
    }
    // epsilon (+ the two median regularisers)
    G_e += -gl * (ep - s.eps_conc) / ee;
    if (b == i_cell) G_e += -gl * (R(1.0) - ee) / R(1e-4);
    if (b == i_empty) G_e += -gl * (R(1.0) - ee) / R(1e-4);
    const R dt_dconc = -(log(ep) + log(ee) - digamma_d(s.eps_conc));
    const R G_conc = -gl * dt_dconc + G_e * R(sgg[1 + b]) / ep;
    const R G_epsenc = G_conc * ep * e.prob + G_dloc * e.ddloc_deps;
    d_eps_out[b] = float(G_epsenc * e.deps_dout);
    // p_y
    const R plp = p_logit_prior_row(e.counts, k);
    const R q1q0 = e.q1 * e.q0;
    const R dt_y = q1q0 * ((logsigmoid_d(plp) - e.lq1) - (logsigmoid_d(-plp) - e.lq0));
    const R dt_plr = -(e.p_y + R(2.0) * s.n_v - R(k.p_logit_prior)) / R(4.0);
    R zq = R(0.0);
    for (int c2 = 0; c2 < zn; ++c2) zq += R(zpart_q[b * zn + c2]);
    const R dt_z = -zq * q1q0;
    R G_py = -gl * (dt_y + dt_plr + dt_z);
    if (d_w) G_py += (-R(d_w[3 * b]) + R(d_w[3 * b + 1]) + R(0.01) * R(d_w[3 * b + 2])) * q1q0;
    d_p_out[b] = float(G_py * e.dpy_dpout);
  }
  if (g.has_rho) {  // droplet-independent digammas, B times each, one per thread of the last warp
    const int t = int(blockDim.x) - 1 - int(threadIdx.x);
    if (t < 3) {
      const double v = double(B) * double(gl) * double(digamma_d(t == 0 ? g.ra + g.rb : t == 1 ? g.ra : g.rb));
      if (t == 0) ga[3] += v, ga[4] += v;
      else if (t == 1) ga[3] -= v;
      else ga[4] -= v;
    }
  }
  R g_pl = R(0.0), g_ps = R(0.0);
  if (threadIdx.x == 0) {  // phi (global)
    const R pc0 = k.phi_conc_prior, pr0 = k.phi_rate_prior;
    const R phi = fmax(R(g_gamma[0]) / g.phi_rate, kFltTiny);
    const R dt_dphi = (pc0 - R(1.0)) / phi - pr0 - (g.phi_conc - R(1.0)) / phi + g.phi_rate;
    const R G_phi = -gl * dt_dphi + (d_alpha ? R(d_alpha[0]) : R(0.0)) * (-R(1.0) / (phi * phi));
    const R dt_dpc = -(log(g.phi_rate) + log(phi) - digamma_d(g.phi_conc));
    const R dt_dpr = -(g.phi_conc / g.phi_rate - phi);
    const R G_pc = -gl * dt_dpc + G_phi * R(sgg[0]) / g.phi_rate;
    const R G_pr = -gl * dt_dpr - G_phi * phi / g.phi_rate;
    const R pl = g.pl, ps = g.psc;
    g_pl = G_pc * R(2.0) * pl / (ps * ps) + G_pr / (ps * ps);
    g_ps = G_pc * (-R(2.0) * pl * pl / (ps * ps * ps)) + G_pr * (-R(2.0) * pl / (ps * ps * ps));
  }
  ga[5] = g_pl, ga[6] = g_ps;
  block_sum_d<7>(ga, red);
  if (threadIdx.x == 0) {  // chain through the constraint transforms (exp; affine o clipped sigmoid)
    if (sg.d_cell_scale) sg.d_cell_scale[0] = float(ga[0] * g.dcs);
    if (sg.d_empty_loc) sg.d_empty_loc[0] = float(ga[1] * g.del);
    if (sg.d_empty_scale) sg.d_empty_scale[0] = float(ga[2] * g.des);
    const R span = R(k.rho_param_max) - R(k.rho_param_min);
    if (sg.rho_alpha) sg.rho_alpha[0] = g.has_rho ? float(ga[3] * span * g.sra * (R(1.0) - g.sra)) : 0.f;
    if (sg.rho_beta) sg.rho_beta[0] = g.has_rho ? float(ga[4] * span * g.srb * (R(1.0) - g.srb)) : 0.f;
    if (sg.phi_loc) sg.phi_loc[0] = float(ga[5] * g.pl);
    if (sg.phi_scale) sg.phi_scale[0] = float(ga[6] * g.psc);
  }
}

// ------------------------------------------------------------------------------------------------------------
// transform_to(constraints.simplex) = torch's SoftmaxTransform: y = softmax(u) over the G genes (pyro.param
// "chi_ambient", model.py:245-249, 486-490).  One CTA (G = 33.5k values: 134 KB, latency-bound); sums in fp64.
constexpr int kSimplexThreads = 1024, kSimplexPer = 40;  // register-cached up to 40 960 genes; longer inputs re-read

__global__ void __launch_bounds__(kSimplexThreads) simplex_forward_kernel(const float* __restrict__ u, int n,
                                                                          float* __restrict__ y, float* __restrict__ y_unit) {
  __shared__ float redf[32];
  __shared__ double red[2 * 32];
  const int t = threadIdx.x;
  const bool cached = n <= kSimplexThreads * kSimplexPer;
  float v[kSimplexPer];
  float mx = -INFINITY;
  if (cached) {
#pragma unroll
    for (int i = 0; i < kSimplexPer; ++i) {
      const int idx = t + i * kSimplexThreads;
      v[i] = idx < n ? u[idx] : -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < kSimplexPer; ++i) mx = fmaxf(mx, v[i]);
  } else {
    for (int i = t; i < n; i += kSimplexThreads) mx = fmaxf(mx, u[i]);
  }
  mx = block_max(mx, redf);
  double acc[2] = {0.0, 0.0};  // sum e, sum e^2
  if (cached) {
#pragma unroll
    for (int i = 0; i < kSimplexPer; ++i) {
      v[i] = expf(v[i] - mx);  // exp(-inf) = 0 for the padding
      acc[0] += double(v[i]);
      acc[1] += double(v[i]) * double(v[i]);
    }
  } else {
    for (int i = t; i < n; i += kSimplexThreads) {
      const double e = double(expf(u[i] - mx));
      acc[0] += e;
      acc[1] += e * e;
    }
  }
  block_sum<double, 2>(acc, red);
  const float inv = float(1.0 / acc[0]), inv_unit = float(1.0 / sqrt(acc[1]));
  if (cached) {
#pragma unroll
    for (int i = 0; i < kSimplexPer; ++i) {
      const int idx = t + i * kSimplexThreads;
      if (idx < n) {
        y[idx] = v[i] * inv;
        if (y_unit) y_unit[idx] = v[i] * inv_unit;
      }
    }
  } else {
    for (int i = t; i < n; i += kSimplexThreads) {
      const float e = expf(u[i] - mx);
      y[i] = e * inv;
      if (y_unit) y_unit[i] = e * inv_unit;
    }
  }
}

// du_k = y_k (gy_k - sum_i gy_i y_i); two sweeps with 8 independent loads in flight per thread
__global__ void __launch_bounds__(kSimplexThreads) simplex_backward_kernel(int n, const float* __restrict__ y,
                                                                           const float* __restrict__ gy,
                                                                           float* __restrict__ du) {
  __shared__ double red[32];
  const int t = threadIdx.x;
  constexpr int U = 8;
  double acc[1] = {0.0};
  for (int i0 = t; i0 < n; i0 += kSimplexThreads * U) {
    float yv[U], gv[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int idx = i0 + i * kSimplexThreads;
      yv[i] = idx < n ? y[idx] : 0.f;
      gv[i] = idx < n ? gy[idx] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) acc[0] += double(gv[i]) * double(yv[i]);
  }
  block_sum<double, 1>(acc, red);
  const float dot = float(acc[0]);
  for (int i0 = t; i0 < n; i0 += kSimplexThreads * U) {
    float yv[U], gv[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int idx = i0 + i * kSimplexThreads;
      yv[i] = idx < n ? y[idx] : 0.f;
      gv[i] = idx < n ? gy[idx] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int idx = i0 + i * kSimplexThreads;
      if (idx < n) du[idx] = yv[i] * (gv[i] - dot);
    }
  }
}

template <typename... P>
static bool aligned16(const P*... p) {  // null pointers (optional operands) count as aligned
  return (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ...);
}

static LatentConsts load_consts(const float* h) {
  LatentConsts k;
  float* dst = reinterpret_cast<float*>(&k);
  for (int i = 0; i < kNumConsts; ++i) dst[i] = h[i];
  return k;
}

static size_t latent_smem(int B, int Z) { return latent_keys_offset(B, Z) + size_t(B) * 12 + 16; }

}  // namespace cb

extern "C" int cb_latent_num_consts(void) { return cb::kNumConsts; }
extern "C" int cb_latent_num_terms(void) { return cb::T_N; }

extern "C" int cb_latent_prepare(const float* feats, const float* p_out, const float* eps_out, int n_rows,
                                 const float* u_d_cell_scale, const float* u_d_empty_loc, const float* u_d_empty_scale,
                                 const float* u_rho_alpha, const float* u_rho_beta, const float* u_phi_loc,
                                 const float* u_phi_scale, const float* consts_h, float* conc_gamma, float* conc_beta,
                                 float* total_beta, void* stream) {
  CB_REQUIRE(n_rows > 0, "cb_latent_prepare: n_rows must be positive");
  cb::ScalarParams sp{u_d_cell_scale, u_d_empty_loc, u_d_empty_scale, u_rho_alpha, u_rho_beta, u_phi_loc, u_phi_scale};
  cb::latent_prepare_kernel<<<(n_rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      feats, p_out, eps_out, n_rows, sp, cb::load_consts(consts_h), conc_gamma, conc_beta, total_beta);
  return cb::check_launch("cb_latent_prepare");
}

extern "C" int cb_latent_beta_sample(float* gamma_draws, int n_rows, float* xx, void* stream) {
  const int n = 1 + 3 * n_rows;
  cb::latent_beta_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gamma_draws, n_rows, xx);
  return cb::check_launch("cb_latent_beta_sample");
}

extern "C" int cb_latent_forward(const float* feats, const float* p_out, const float* eps_out, const float* z_loc,
                                 const float* z_log_scale, long long ldz, int n_rows, int z_dim, const float* noise_z,
                                 const float* noise_row, const float* gamma_draws, const float* xx,
                                 const float* u_d_cell_scale, const float* u_d_empty_loc, const float* u_d_empty_scale,
                                 const float* u_rho_alpha, const float* u_rho_beta, const float* u_phi_loc,
                                 const float* u_phi_scale, const float* consts_h, float* z_out, float* coef, float* alpha,
                                 float* w, float* terms, float* loss_local, float* enc_out, int* med_idx,
                                 void* stream) {
  CB_REQUIRE(n_rows > 0 && n_rows <= 2048 && z_dim > 0 && z_dim <= 256, "cb_latent_forward: n_rows %d / z_dim %d out of range", n_rows, z_dim);
  cb::ScalarParams sp{u_d_cell_scale, u_d_empty_loc, u_d_empty_scale, u_rho_alpha, u_rho_beta, u_phi_loc, u_phi_scale};
  const int vec = z_dim % 4 == 0 && ldz % 4 == 0 && cb::aligned16(z_loc, z_log_scale, noise_z, z_out);
  cb::latent_forward_kernel<<<1, cb::kLatentThreads, cb::latent_smem(n_rows, z_dim), (cudaStream_t)stream>>>(
      feats, p_out, eps_out, z_loc, z_log_scale, ldz, n_rows, z_dim, noise_z, noise_row, gamma_draws, xx, sp,
      cb::load_consts(consts_h), z_out, coef, alpha, w, terms, loss_local, enc_out, med_idx, vec);
  return cb::check_launch("cb_latent_forward");
}

extern "C" int cb_latent_backward(const float* feats, const float* p_out, const float* eps_out, const float* z_loc,
                                  const float* z_log_scale, long long ldz, int n_rows, int z_dim, const float* noise_z,
                                  const float* noise_row, const float* gamma_draws, const float* xx,
                                  const float* gamma_grad, const float* dirichlet_grad, const int* med_idx,
                                  const float* u_d_cell_scale, const float* u_d_empty_loc, const float* u_d_empty_scale,
                                  const float* u_rho_alpha, const float* u_rho_beta, const float* u_phi_loc,
                                  const float* u_phi_scale, const float* consts_h, const float* d_z, const float* d_coef,
                                  const float* d_alpha, const float* d_w, const float* g_local, float* d_p_out,
                                  float* d_eps_out, float* d_z_loc, float* d_z_log_scale, long long lddz,
                                  float* g_d_cell_scale, float* g_d_empty_loc, float* g_d_empty_scale, float* g_rho_alpha,
                                  float* g_rho_beta, float* g_phi_loc, float* g_phi_scale, void* stream) {
  CB_REQUIRE(n_rows > 0 && n_rows <= 2048 && z_dim > 0 && z_dim <= 256, "cb_latent_backward: n_rows %d / z_dim %d out of range", n_rows, z_dim);
  cb::ScalarParams sp{u_d_cell_scale, u_d_empty_loc, u_d_empty_scale, u_rho_alpha, u_rho_beta, u_phi_loc, u_phi_scale};
  cb::ScalarGrads sg{g_d_cell_scale, g_d_empty_loc, g_d_empty_scale, g_rho_alpha, g_rho_beta, g_phi_loc, g_phi_scale};
  const int vec = z_dim % 4 == 0 && ldz % 4 == 0 && lddz % 4 == 0 &&
                  cb::aligned16(z_loc, z_log_scale, noise_z, d_z, d_z_loc, d_z_log_scale);
  cb::latent_backward_kernel<<<1, cb::kLatentThreads, cb::latent_smem(n_rows, z_dim), (cudaStream_t)stream>>>(
      feats, p_out, eps_out, z_loc, z_log_scale, ldz, n_rows, z_dim, noise_z, noise_row, gamma_draws, xx, gamma_grad,
      dirichlet_grad, med_idx, sp, cb::load_consts(consts_h), d_z, d_coef, d_alpha, d_w, g_local, d_p_out, d_eps_out, d_z_loc,
      d_z_log_scale, lddz, sg, vec);
  return cb::check_launch("cb_latent_backward");
}

extern "C" int cb_simplex_forward(const float* u, int n, float* y, float* y_unit, void* stream) {
  CB_REQUIRE(n >= 1, "cb_simplex_forward: empty parameter");
  cb::simplex_forward_kernel<<<1, cb::kSimplexThreads, 0, (cudaStream_t)stream>>>(u, n, y, y_unit);
  return cb::check_launch("cb_simplex_forward");
}

extern "C" int cb_simplex_backward(int n, const float* y, const float* grad_y, float* grad_u, void* stream) {
  CB_REQUIRE(n >= 1, "cb_simplex_backward: empty parameter");
  cb::simplex_backward_kernel<<<1, cb::kSimplexThreads, 0, (cudaStream_t)stream>>>(n, y, grad_y, grad_u);
  return cb::check_launch("cb_simplex_backward");
}
