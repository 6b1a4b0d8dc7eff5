// K9/K10: posterior noise-count log-probability over the NON-ZERO counts only, all latent samples
// accumulated on chip, emitted directly as the sparse COO the reference builds.
//
// Replaces (reference cellbender/remove_background/posterior.py):
//   _log_prob_noise_count_tensor  :784-859   dense [128, G, 20] fp32 temporaries (343 MB each), x 20 samples
//   noise_log_pdf                 :669-735   logsumexp normalisation + running log-mean over samples
//   _get_cell_noise_count_posterior_coo :528-576  exp / two masked writes / nonzero / CPU index translation
//   sparse_utils.dense_to_sparse_op_torch   sparse_utils.py:10-33;  IndexConverter.get_m_indices  posterior.py:1554-1575
// The reference evaluates every (cell, gene, c) and then throws away x == 0 (92-95 % of the work).
//
// Data layout (r02): the decoder output of a chunk of cells is produced TRANSPOSED and cell-major,
//   logits_t[g, b * S + s] = W_dec[g, :] . h[b * S + s, :]            (cb_gemm_tf32 with A = W_dec, B = h; no bias)
// so that the S latent samples of one (cell, gene) are S consecutive floats (80 bytes at S = 20: 3 sectors instead of
// the 20 scattered sectors of a sample-major [S * cells, G] layout).  Four lanes share one stored count: lane j walks
// samples j, j + 4, ... (each load instruction of the quad reads 16 contiguous bytes), the quad's partial window pmfs are
// combined with two shuffle steps.  The softmax denominators come from cb_col_logsumexp over the same matrix.
//
// Quirks reproduced on purpose:
//   * the window size n = min(n_counts_max, max count of the reference minibatch) is passed per cell;
//   * per-sample windows may start at different `low`; the reference averages them index-wise and reports
//     the LAST sample's low as the offset (posterior.py:706-731) -- so do we.
#include "common.cuh"
#include "posterior_math.cuh"

namespace cb {

constexpr int kPostThreads = 256;  // 64 stored counts per CTA
// exp as one FMUL + one ex2.approx.ftz (relative error 2^-22)
__device__ __forceinline__ float exp_fast(float v) { return pm_exp2(v * 1.4426950408889634f); }

// logp_out [E, NMAX] : log of the sample-mean pmf;  low_out [E];  keep_cnt [E] = #(c : logp >= thresh);
// entry_m [E] = barcode_all[cell] * total_n_genes + gene_all[g].  E = entry_start[n_cells] stored counts of the chunk,
// in (cell, gene) order.
template <int NMAX>
__global__ void __launch_bounds__(kPostThreads) posterior_window_kernel(
    const long long* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
    const long long* __restrict__ cell_rows, int n_cells, int n_samples, const float* __restrict__ logits_t,
    long long ldt, const float* __restrict__ dec_bias, const float* __restrict__ lse, const float* __restrict__ chi_amb,
    const float* __restrict__ chi_bar, const float* __restrict__ a_mu, const float* __restrict__ c_a,
    const float* __restrict__ c_b, const float* __restrict__ alpha, const int* __restrict__ n_window,
    const long long* __restrict__ entry_start, const long long* __restrict__ barcode_all,
    const long long* __restrict__ gene_all, long long total_n_genes, float log_thresh, float lambda_multiplier,
    const long long* __restrict__ out_offset, float* __restrict__ logp_out, int* __restrict__ low_out,
    int* __restrict__ keep_cnt, long long* __restrict__ entry_m) {
  // the grid may be sized for an upper bound of the chunk's stored counts (CUDA-graph replays), or capped at a few CTAs
  // per SM so that the kernel shares the GPU with the decoder product of the next chunk: tiles are strided over the CTAs
  const long long n_entries = entry_start[n_cells];
  const long long o0 = out_offset ? out_offset[0] : 0;  // first output row of this chunk (device scalar)
  logp_out += o0 * NMAX, low_out += o0, keep_cnt += o0, entry_m += o0;
  const int sub = threadIdx.x & 3;
  for (long long tile = blockIdx.x; tile * (kPostThreads / 4) < n_entries; tile += gridDim.x) {
  const long long e = tile * (long long)(kPostThreads / 4) + (threadIdx.x >> 2);
  const bool active = e < n_entries;
  // cell of this entry: last b with entry_start[b] <= e
  int b = 0;
  if (active) {
    int lo = 0, hi = n_cells;  // invariant: entry_start[lo] <= e < entry_start[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (entry_start[mid] <= e) lo = mid;
      else hi = mid;
    }
    b = lo;
  }
  float acc[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) acc[k] = 0.f;
  int low = 0, n = 0, g = 0;
  float x = 0.f;
  if (active) {
    const long long j = indptr[cell_rows[b]] + (e - entry_start[b]);
    x = data[j];
    g = indices[j];
    n = n_window[b] < NMAX ? n_window[b] : NMAX;
  }
  // live window terms of this entry are at most min(n, x + 1) whatever the sample's window start: the warp's maximum bounds
  // the unrolled term loops for all eight entries of the warp (most stored counts are 1..3: a few terms instead of 20)
  const int my_bound = (active && x > 0.f && n > 0) ? min(n, int(fminf(x, float(NMAX))) + 1) : 0;
  const int live_bound = __reduce_max_sync(0xffffffffu, my_bound);
  if (active) {
    if (x > 0.f && n > 0) {
      const float ca = chi_amb[g], cb_ = chi_bar[g], bg = dec_bias ? dec_bias[g] : 0.f;
      const long long base = (long long)b * n_samples;
      const float* lrow = logits_t + (long long)g * ldt + base;
      for (int smp = sub; smp < n_samples; smp += 4) {
        const float chi = exp_fast(lrow[smp] + bg - lse[base + smp]);
        // posterior.py:708-710 safeguards
        const float mu = a_mu[base + smp] * chi + 1e-30f;
        const float lam = (c_a[base + smp] * ca + c_b[base + smp] * cb_) * lambda_multiplier + 1e-30f;
        const float al = alpha[smp] + 1e-30f;
        const int lw = noise_window_pmf_accumulate<NMAX>(x, mu, lam, al, n, acc, live_bound);
        if (smp == n_samples - 1) low = lw;
      }
    }
  }
  // combine the quad (all 32 lanes take part in the shuffles)
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
  }
  low += __shfl_xor_sync(0xffffffffu, low, 1);  // exactly one lane of the quad holds the last sample's low
  low += __shfl_xor_sync(0xffffffffu, low, 2);
  const float inv_s = 1.0f / float(n_samples);
  int cnt = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      if ((k & 3) == sub) {  // lane j of the quad finishes columns j, j + 4, ...: 16 contiguous bytes per store instruction
        const float lp = (acc[k] > 0.f) ? 0.6931471805599453f * pm_log2(acc[k] * inv_s) : -INFINITY;  // |error| < 2e-7
        logp_out[e * NMAX + k] = lp;
        cnt += (k < n && lp >= log_thresh) ? 1 : 0;
      }
    }
  }
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
  if (active && sub == 0) {
    low_out[e] = low;
    keep_cnt[e] = cnt;
    entry_m[e] = barcode_all[b] * total_n_genes + gene_all[g];
  }
  }  // tile loop
}

// Writes the COO triplets in (cell, gene, c) order -- the order torch.nonzero gives the reference.
template <int NMAX>
__global__ void __launch_bounds__(256) posterior_compact_kernel(
    long long n_entries, const long long* __restrict__ entry_m, const float* __restrict__ logp, const int* __restrict__ low,
    const int* __restrict__ keep_cnt, const long long* __restrict__ keep_pos, const long long* __restrict__ off_pos,
    float log_thresh, long long* __restrict__ coo_m, int* __restrict__ coo_c, float* __restrict__ coo_lp,
    long long* __restrict__ off_m, int* __restrict__ off_v) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  if (keep_cnt[i] == 0) return;
  long long o = keep_pos[i];
  const long long m = entry_m[i];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    const float lp = logp[i * NMAX + k];
    if (lp >= log_thresh) {
      coo_m[o] = m;
      coo_c[o] = k;
      coo_lp[o] = lp;
      ++o;
    }
  }
  if (low[i] != 0) {
    const long long q = off_pos[i];
    off_m[q] = m;
    off_v[q] = low[i];
  }
}

// has_off[i] = keep_cnt[i] > 0 && low[i] != 0  (entries that contribute an (m, offset) pair)
__global__ void __launch_bounds__(256) posterior_offset_flags_kernel(long long n, const int* __restrict__ keep_cnt,
                                                                     const int* __restrict__ low, int* __restrict__ flag) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (keep_cnt[i] > 0 && low[i] != 0) ? 1 : 0;
}

// ---- column logsumexp of a [R, ld] matrix with an optional per-row additive term: out[c] = log sum_r exp(in[r, c] + add[r]) ----
// pass 1: a CTA owns 128 columns x one of `n_split` row ranges, online (max, sum) per thread; pass 2 merges the ranges.
// exp as ONE ex2.approx.ftz (relative error 2^-22 per term of a 33 538-term sum; arguments <= 0, a flushed term is < 1e-38
// of the largest): with expf the pass was issue-bound at 0.60 of the HBM rate (ncu r02i: issue 72 %, alu 48 %).
constexpr int kColLseThreads = 128;
__global__ void __launch_bounds__(kColLseThreads) col_lse_partial_kernel(const float* __restrict__ in, int R, int C, long long ld,
                                                                         const float* __restrict__ add, int rows_per_split,
                                                                         float* __restrict__ part_max, float* __restrict__ part_sum) {
  const int c = blockIdx.x * kColLseThreads + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per_split, r1 = min(R, r0 + rows_per_split);
  float mx = -INFINITY, sm = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {  // four independent loads in flight per thread
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = in[(long long)(r + u) * ld + c] + (add ? add[r + u] : 0.f);
    const float m4 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    if (m4 > mx) {
      sm *= exp_fast(mx - m4);  // exp(-inf) = 0 on the first group
      mx = m4;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) sm += exp_fast(v[u] - mx);
  }
  for (; r < r1; ++r) {
    const float v = in[(long long)r * ld + c] + (add ? add[r] : 0.f);
    if (v > mx) {
      sm *= exp_fast(mx - v);
      mx = v;
    }
    sm += exp_fast(v - mx);
  }
  part_max[(long long)blockIdx.y * C + c] = mx;
  part_sum[(long long)blockIdx.y * C + c] = sm;
}

__global__ void __launch_bounds__(256) col_lse_final_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                                            int n_split, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mx = -INFINITY;
  for (int s = 0; s < n_split; ++s) mx = fmaxf(mx, part_max[(long long)s * C + c]);
  float sm = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float pm = part_max[(long long)s * C + c];
    if (pm > -INFINITY) sm += part_sum[(long long)s * C + c] * expf(pm - mx);
  }
  out[c] = mx + logf(sm);
}

}  // namespace cb

extern "C" int cb_posterior_noise_window(const long long* indptr, const int* indices, const float* data,
                                         const long long* cell_rows, int n_cells, int n_genes, int n_samples,
                                         const float* logits_t, long long ldt, const float* dec_bias, const float* lse,
                                         const float* chi_ambient, const float* chi_bar, const float* a_mu, const float* c_a,
                                         const float* c_b, const float* alpha, const int* n_window,
                                         const long long* entry_start, long long max_entries, const long long* barcode_all,
                                         const long long* gene_all, long long total_n_genes, int n_counts_max,
                                         float log_thresh, float lambda_multiplier, const long long* out_offset,
                                         float* logp_out, int* low_out, int* keep_cnt, long long* entry_m, int ctas_per_sm,
                                         void* stream) {
  if (n_cells == 0 || max_entries == 0) return 0;
  const long long n_entries = max_entries;
  CB_REQUIRE(n_cells > 0 && n_genes > 0 && n_samples > 0 && n_entries > 0, "cb_posterior_noise_window: bad shape");
  CB_REQUIRE(ldt >= (long long)n_cells * n_samples, "cb_posterior_noise_window: ldt=%lld < n_cells * n_samples", ldt);
  CB_REQUIRE(n_counts_max > 0 && n_counts_max <= 64, "cb_posterior_noise_window: n_counts_max=%d not in [1, 64]", n_counts_max);
  long long tiles = (n_entries + cb::kPostThreads / 4 - 1) / (cb::kPostThreads / 4);
  if (ctas_per_sm > 0 && tiles > (long long)ctas_per_sm * cb::sm_count()) tiles = (long long)ctas_per_sm * cb::sm_count();
  const unsigned grid = unsigned(tiles);
#define CB_LAUNCH_POST(NMAX)                                                                                          \
  cb::posterior_window_kernel<NMAX><<<grid, cb::kPostThreads, 0, (cudaStream_t)stream>>>(                              \
      indptr, indices, data, cell_rows, n_cells, n_samples, logits_t, ldt, dec_bias, lse, chi_ambient, chi_bar, a_mu, \
      c_a, c_b, alpha, n_window, entry_start, barcode_all, gene_all, total_n_genes, log_thresh, lambda_multiplier,    \
      out_offset, logp_out, low_out, keep_cnt, entry_m)
  if (n_counts_max <= 20) CB_LAUNCH_POST(20);
  else if (n_counts_max <= 32) CB_LAUNCH_POST(32);
  else CB_LAUNCH_POST(64);
#undef CB_LAUNCH_POST
  return cb::check_launch("cb_posterior_noise_window");
}

// row stride of logp_out for a given n_counts_max (20, 32 or 64)
extern "C" int cb_posterior_window_stride(int n_counts_max) { return n_counts_max <= 20 ? 20 : (n_counts_max <= 32 ? 32 : 64); }

extern "C" int cb_posterior_offset_flags(long long n_entries, const int* keep_cnt, const int* low, int* flag, void* stream) {
  if (n_entries == 0) return 0;
  cb::posterior_offset_flags_kernel<<<unsigned((n_entries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_entries, keep_cnt, low,
                                                                                                      flag);
  return cb::check_launch("cb_posterior_offset_flags");
}

extern "C" int cb_posterior_compact(long long n_entries, int n_counts_max, const long long* entry_m, const float* logp,
                                    const int* low, const int* keep_cnt, const long long* keep_pos,
                                    const long long* off_pos, float log_thresh, long long* coo_m, int* coo_c,
                                    float* coo_lp, long long* off_m, int* off_v, void* stream) {
  if (n_entries == 0) return 0;
  const int stride = cb_posterior_window_stride(n_counts_max);
  const unsigned grid = unsigned((n_entries + 255) / 256);
#define CB_LAUNCH_COMPACT(NMAX)                                                                                     \
  cb::posterior_compact_kernel<NMAX><<<grid, 256, 0, (cudaStream_t)stream>>>(n_entries, entry_m, logp, low, keep_cnt, \
                                                                             keep_pos, off_pos, log_thresh, coo_m,  \
                                                                             coo_c, coo_lp, off_m, off_v)
  if (stride == 20) CB_LAUNCH_COMPACT(20);
  else if (stride == 32) CB_LAUNCH_COMPACT(32);
  else CB_LAUNCH_COMPACT(64);
#undef CB_LAUNCH_COMPACT
  return cb::check_launch("cb_posterior_compact");
}

// scratch floats cb_col_logsumexp needs for an [n_rows, n_cols] matrix
extern "C" long long cb_col_logsumexp_scratch_elems(int n_rows, int n_cols) {
  const int n_split = n_rows >= 64 * 64 ? 64 : (n_rows + 63) / 64;
  return 2LL * (n_split < 1 ? 1 : n_split) * n_cols;
}

extern "C" int cb_col_logsumexp(const float* in, int n_rows, int n_cols, long long ld, const float* row_add, float* scratch,
                                float* out, void* stream) {
  if (n_cols == 0) return 0;
  CB_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols, "cb_col_logsumexp: bad shape [%d, %d] ld=%lld", n_rows, n_cols, ld);
  int n_split = n_rows >= 64 * 64 ? 64 : (n_rows + 63) / 64;
  if (n_split < 1) n_split = 1;
  const int rows_per_split = (n_rows + n_split - 1) / n_split;
  float* part_max = scratch;
  float* part_sum = scratch + (long long)n_split * n_cols;
  dim3 grid((n_cols + cb::kColLseThreads - 1) / cb::kColLseThreads, n_split);
  cb::col_lse_partial_kernel<<<grid, cb::kColLseThreads, 0, (cudaStream_t)stream>>>(in, n_rows, n_cols, ld, row_add, rows_per_split,
                                                                                  part_max, part_sum);
  cb::col_lse_final_kernel<<<(n_cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part_max, part_sum, n_split, n_cols, out);
  return cb::check_launch("cb_col_logsumexp");
}
