// K9/K10: posterior noise-count log-probability over the NON-ZERO counts only, all latent samples
// accumulated on chip, emitted directly as the sparse COO the reference builds.
//
// Replaces (reference cellbender/remove_background/posterior.py):
//   _log_prob_noise_count_tensor  :784-859   dense [128, G, 20] fp32 temporaries (343 MB each), x 20 samples
//   noise_log_pdf                 :669-735   logsumexp normalisation + running log-mean over samples
//   _get_cell_noise_count_posterior_coo :528-576  exp / two masked writes / nonzero / CPU index translation
//   sparse_utils.dense_to_sparse_op_torch   sparse_utils.py:10-33
// The reference evaluates every (cell, gene, c) and then throws away x == 0 (92-95 % of the work); here one
// thread owns one stored count and walks the S samples with the window pmf in registers.
//
// Quirks reproduced on purpose:
//   * the window size n = min(n_counts_max, max count of the reference minibatch) is passed per cell;
//   * per-sample windows may start at different `low`; the reference averages them index-wise and reports
//     the LAST sample's low as the offset (posterior.py:706-731) -- so do we.
#include "common.cuh"
#include "posterior_math.cuh"

namespace cb {

constexpr int kPostThreads = 128;

// acc_out [E, NMAX] : log of the sample-mean pmf;  low_out [E];  keep_cnt [E] = #(c : logp >= thresh & prob != 0)
template <int NMAX>
__global__ void __launch_bounds__(kPostThreads) posterior_window_kernel(
    const long long* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ data,
    const long long* __restrict__ cell_rows, int n_cells, int n_samples, const float* __restrict__ logits,
    long long ldl, const float* __restrict__ lse, const float* __restrict__ chi_amb, const float* __restrict__ chi_bar,
    const float* __restrict__ a_mu, const float* __restrict__ c_a, const float* __restrict__ c_b,
    const float* __restrict__ alpha, const int* __restrict__ n_window, const long long* __restrict__ entry_start,
    float log_thresh, float lambda_multiplier, float* __restrict__ logp_out, int* __restrict__ low_out,
    int* __restrict__ keep_cnt) {
  const int b = blockIdx.x;
  const long long r = cell_rows[b];
  const long long s = indptr[r], e = indptr[r + 1];
  const long long out0 = entry_start[b];
  const int n = n_window[b] < NMAX ? n_window[b] : NMAX;
  const float inv_s = 1.0f / float(n_samples);
  for (long long j = s + threadIdx.x; j < e; j += blockDim.x) {
    const long long eo = out0 + (j - s);
    const float x = data[j];
    int low = 0, cnt = 0;
    float acc[NMAX];
#pragma unroll
    for (int k = 0; k < NMAX; ++k) acc[k] = 0.f;
    if (x > 0.f && n > 0) {
      const int g = indices[j];
      const float ca = chi_amb[g], cb_ = chi_bar[g];
      for (int smp = 0; smp < n_samples; ++smp) {
        const long long row = (long long)smp * n_cells + b;
        const float chi = expf(logits[row * ldl + g] - lse[row]);
        // posterior.py:708-710 safeguards
        const float mu = a_mu[row] * chi + 1e-30f;
        const float lam = (c_a[row] * ca + c_b[row] * cb_) * lambda_multiplier + 1e-30f;
        const float al = alpha[smp] + 1e-30f;
        float p[NMAX];
        low = noise_window_pmf<NMAX>(x, mu, lam, al, n, p);
#pragma unroll
        for (int k = 0; k < NMAX; ++k) acc[k] += p[k];
      }
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      const float lp = (acc[k] > 0.f) ? logf(acc[k] * inv_s) : -INFINITY;
      logp_out[eo * NMAX + k] = lp;
      cnt += (k < n && lp >= log_thresh) ? 1 : 0;
    }
    low_out[eo] = low;
    keep_cnt[eo] = cnt;
  }
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

// m index of every stored count of the selected cells: m = barcode_all[cell] * G_all + gene_all[g]  (IndexConverter)
__global__ void __launch_bounds__(kPostThreads) posterior_entry_index_kernel(
    const long long* __restrict__ indptr, const int* __restrict__ indices, const long long* __restrict__ cell_rows,
    const long long* __restrict__ barcode_all, const long long* __restrict__ gene_all, long long total_n_genes,
    const long long* __restrict__ entry_start, long long* __restrict__ entry_m) {
  const int b = blockIdx.x;
  const long long r = cell_rows[b];
  const long long s = indptr[r], e = indptr[r + 1];
  const long long base = barcode_all[b] * total_n_genes;
  for (long long j = s + threadIdx.x; j < e; j += blockDim.x) entry_m[entry_start[b] + (j - s)] = base + gene_all[indices[j]];
}

}  // namespace cb

extern "C" int cb_posterior_noise_window(const long long* indptr, const int* indices, const float* data,
                                         const long long* cell_rows, int n_cells, int n_genes, int n_samples,
                                         const float* logits, long long ldl, const float* lse, const float* chi_ambient,
                                         const float* chi_bar, const float* a_mu, const float* c_a, const float* c_b,
                                         const float* alpha, const int* n_window, const long long* entry_start,
                                         int n_counts_max, float log_thresh, float lambda_multiplier, float* logp_out,
                                         int* low_out, int* keep_cnt, void* stream) {
  if (n_cells == 0) return 0;
  CB_REQUIRE(n_cells > 0 && n_genes > 0 && n_samples > 0, "cb_posterior_noise_window: bad shape");
  CB_REQUIRE(n_counts_max > 0 && n_counts_max <= 64, "cb_posterior_noise_window: n_counts_max=%d not in [1, 64]", n_counts_max);
#define CB_LAUNCH_POST(NMAX)                                                                                         \
  cb::posterior_window_kernel<NMAX><<<n_cells, cb::kPostThreads, 0, (cudaStream_t)stream>>>(                          \
      indptr, indices, data, cell_rows, n_cells, n_samples, logits, ldl, lse, chi_ambient, chi_bar, a_mu, c_a, c_b, \
      alpha, n_window, entry_start, log_thresh, lambda_multiplier, logp_out, low_out, keep_cnt)
  if (n_counts_max <= 20) CB_LAUNCH_POST(20);
  else if (n_counts_max <= 32) CB_LAUNCH_POST(32);
  else CB_LAUNCH_POST(64);
#undef CB_LAUNCH_POST
  return cb::check_launch("cb_posterior_noise_window");
}

// row stride of logp_out for a given n_counts_max (20, 32 or 64)
extern "C" int cb_posterior_window_stride(int n_counts_max) { return n_counts_max <= 20 ? 20 : (n_counts_max <= 32 ? 32 : 64); }

extern "C" int cb_posterior_entry_index(const long long* indptr, const int* indices, const long long* cell_rows,
                                        int n_cells, const long long* barcode_all, const long long* gene_all,
                                        long long total_n_genes, const long long* entry_start, long long* entry_m,
                                        void* stream) {
  if (n_cells == 0) return 0;
  cb::posterior_entry_index_kernel<<<n_cells, cb::kPostThreads, 0, (cudaStream_t)stream>>>(
      indptr, indices, cell_rows, barcode_all, gene_all, total_n_genes, entry_start, entry_m);
  return cb::check_launch("cb_posterior_entry_index");
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
