// K11: noise-count estimators as segmented reductions over the (m, c)-sorted posterior COO.
//
// Replaces (reference cellbender/remove_background/estimation.py):
//   apply_function_dense_chunks / chunked_iterator  :705-793   (np.unique + pd.isin + densify to float64, 1 GB chunks)
//   MAP / Mean / ThresholdCDF / SingleSample        :89-204
//   _estimation_array_to_csr                        :207-235
//   MultipleChoiceKnapsack._chunk_estimate_noise    :552-702   (pandas apply/sort/diff/groupby-nsmallest, ~0.18 M entries/s)
// A "segment" is the run of COO entries that share one m = n * G_all + g; entries inside a segment are sorted
// by c.  Everything here is integer / comparison work on <= ~20 entries per segment: HBM-bound, one thread per
// segment or per entry, no atomics on floating point (results are order-independent and bit-reproducible).
#include "common.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------------------
// exclusive prefix sum over int32 counts -> int64 positions (three-kernel scan, deterministic)
// ---------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const int* __restrict__ in, long long n,
                                                                      long long* __restrict__ tile_sums) {
  __shared__ long long red[32];
  const long long base = (long long)blockIdx.x * kScanTile;
  long long s = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const long long i = base + (long long)k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long v = (threadIdx.x < kScanThreads / 32) ? red[threadIdx.x] : 0;
    v = warp_sum(v);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = v;
  }
}

// single block: exclusive scan of tile sums in place; total written to tile_sums[n_tiles]
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(long long* __restrict__ tile_sums, long long n_tiles) {
  __shared__ long long buf[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n_tiles; base += 1024) {
    const long long i = base + threadIdx.x;
    const long long v = (i < n_tiles) ? tile_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      long long t = (threadIdx.x >= off) ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_tiles) tile_sums[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ in, long long n,
                                                                  const long long* __restrict__ tile_offsets,
                                                                  long long* __restrict__ out) {
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
  __shared__ long long wsum[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  long long local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    local += v[k];
  }
  // exclusive scan of `local` across the block
  long long incl = local;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  long long woff = 0;
  for (int i = 0; i < wid; ++i) woff += wsum[i];
  long long run = tile_offsets[blockIdx.x] + woff + incl - local;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

// ---------------------------------------------------------------------------------------------------------
// segments
// ---------------------------------------------------------------------------------------------------------
// head[i] = 1 where m[i] != m[i-1]; also checks (m, c) sortedness -> *unsorted set to 1 if violated
__global__ void __launch_bounds__(256) segment_heads_kernel(const long long* __restrict__ m, const int* __restrict__ c,
                                                            long long n, int* __restrict__ head,
                                                            int* __restrict__ unsorted) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) {
    head[0] = 1;
    return;
  }
  const long long a = m[i - 1], b = m[i];
  head[i] = (a != b) ? 1 : 0;
  if (a > b || (a == b && c[i - 1] >= c[i])) atomicOr(unsorted, 1);
}

// seg_start[seg_id] = i for heads; seg_of[i] = segment index of entry i (pos = exclusive scan of head)
__global__ void __launch_bounds__(256) segment_starts_kernel(const int* __restrict__ head, const long long* __restrict__ pos,
                                                             long long n, long long* __restrict__ seg_start,
                                                             int* __restrict__ seg_of) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long sidx = pos[i] + head[i] - 1;  // inclusive count - 1
  seg_of[i] = int(sidx);
  if (head[i]) seg_start[sidx] = i;
}

__device__ __forceinline__ int lookup_offset(const long long* __restrict__ off_m, const int* __restrict__ off_v,
                                             long long n_off, long long key) {
  long long lo = 0, hi = n_off;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (off_m[mid] < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n_off && off_m[lo] == key) ? off_v[lo] : 0;
}

// philox-free counter hash -> uniform in (0,1): splitmix64 (deterministic given seed + segment index)
__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

enum { kModeMap = 0, kModeMean = 1, kModeCdf = 2, kModeSample = 3 };

// One thread per segment.  out_i (int32) for MAP/CDF/Sample, out_f (fp32) for Mean; offsets added.
// map_pos (optional): index of the MAP entry inside the COO (for MCKP).
__global__ void __launch_bounds__(256) segment_estimate_kernel(
    int mode, const long long* __restrict__ m, const int* __restrict__ c, const float* __restrict__ lp,
    const long long* __restrict__ seg_start, long long n_seg, long long n_entries, const long long* __restrict__ off_m,
    const int* __restrict__ off_v, long long n_off, double q, int n_cols, unsigned long long seed,
    long long* __restrict__ seg_m, int* __restrict__ out_i, float* __restrict__ out_f, long long* __restrict__ map_pos) {
  const long long sidx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (sidx >= n_seg) return;
  const long long a = seg_start[sidx];
  const long long b = (sidx + 1 < n_seg) ? seg_start[sidx + 1] : n_entries;
  const long long key = m[a];
  seg_m[sidx] = key;
  const int off = lookup_offset(off_m, off_v, n_off, key);
  if (mode == kModeMap) {
    long long best = a;
    float bv = lp[a];
    for (long long i = a + 1; i < b; ++i) {
      const float v = lp[i];
      if (v > bv) bv = v, best = i;  // strict: first maximum wins (torch.argmax / np.argmax)
    }
    out_i[sidx] = c[best] + off;
    if (map_pos != nullptr) map_pos[sidx] = best;
  } else if (mode == kModeMean) {
    double acc = 0.0;
    for (long long i = a; i < b; ++i) acc += exp(double(lp[i])) * double(c[i]);
    out_f[sidx] = float(acc + double(off));
  } else if (mode == kModeCdf) {
    // (exp(logp).cumsum(-1) <= q).sum(-1) over the dense columns 0..n_cols-1 (missing entries add 0)
    double cum = 0.0;
    int cnt = 0, prev = -1;
    for (long long i = a; i < b; ++i) {
      const int ci = c[i];
      if (cum <= q) cnt += ci - prev - 1;
      cum += exp(double(lp[i]));
      if (cum <= q) cnt += 1;
      prev = ci;
    }
    if (cum <= q) cnt += n_cols - 1 - prev;
    out_i[sidx] = cnt + off;
  } else {  // Categorical(logits = lp) sample by inverse CDF
    double mx = -INFINITY, z = 0.0;
    for (long long i = a; i < b; ++i) mx = fmax(mx, double(lp[i]));
    for (long long i = a; i < b; ++i) z += exp(double(lp[i]) - mx);
    const double u = uniform01(seed, (unsigned long long)sidx) * z;
    double cum = 0.0;
    long long pick = b - 1;
    for (long long i = a; i < b; ++i) {
      cum += exp(double(lp[i]) - mx);
      if (u < cum) {
        pick = i;
        break;
      }
    }
    // never pick a zero-probability (-inf) entry
    while (pick > a && !(lp[pick] > -INFINITY)) --pick;
    out_i[sidx] = c[pick] + off;
  }
}

// ---------------------------------------------------------------------------------------------------------
// PRq posterior regularisation (reference posterior.py:1002-1225): per m-segment
//   target = log(mean + alpha * std) of the noise-count posterior (dense column index c as the count, fp32 c times
//            fp64 probabilities, as _log_mean_plus_alpha_std evaluates it)
//   beta   = bisection on [-100, 100] (torch_binary_search, posterior.py:1654-1746: fp32 bracket, midpoint, stop when
//            |violation| < tol, lower bound moves when violation < -tol, upper when > tol, at most max_iter rounds) of
//            violation(beta) = LSE_c(log n_c + lp_c + beta n_c) - LSE_c(lp_c + beta n_c) - target,  n_c = c + offset_m
//            (n_c, log n_c and beta n_c in fp32, the sums in fp64: the reference's dtype flow)
//   output = lp_c + beta n_c - LSE_c(lp_c + beta n_c);  keep[entry] = exp(output) != 0 in fp64.
// The reference iterates all segments of a chunk until ALL have converged; a converged segment's bracket no longer
// moves, so per-segment loops give the same beta.  One thread per segment, at most ~20 entries each.
__global__ void __launch_bounds__(256) prq_regularize_kernel(
    const long long* __restrict__ m, const int* __restrict__ c, const float* __restrict__ lp,
    const long long* __restrict__ seg_start, long long n_seg, long long n_entries, const long long* __restrict__ off_m,
    const int* __restrict__ off_v, long long n_off, double alpha, double tol, int max_iter, double* __restrict__ log_target,
    float* __restrict__ beta_out, double* __restrict__ lp_out, int* __restrict__ keep) {
  const long long sidx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (sidx >= n_seg) return;
  const long long a = seg_start[sidx];
  const long long b = (sidx + 1 < n_seg) ? seg_start[sidx + 1] : n_entries;
  const int off = lookup_offset(off_m, off_v, n_off, m[a]);
  double mean = 0.0;
  for (long long i = a; i < b; ++i) mean += double(float(c[i])) * exp(double(lp[i]));
  double var = 0.0;
  for (long long i = a; i < b; ++i) {
    const double d = double(float(c[i])) - mean;
    var += d * d * exp(double(lp[i]));
  }
  const double target = log(mean + alpha * sqrt(var));
  if (log_target != nullptr) log_target[sidx] = target;

  float lo = -100.0f, hi = 100.0f, beta = 0.0f;
  for (int it = 0; it < max_iter; ++it) {
    beta = (lo + hi) * 0.5f;
    double mx_n = -INFINITY, mx_d = -INFINITY;
    for (long long i = a; i < b; ++i) {
      const float n = float(c[i] + off);
      const double td = double(lp[i]) + double(beta * n), tn = double(logf(n)) + td;
      mx_n = fmax(mx_n, tn), mx_d = fmax(mx_d, td);
    }
    double s_n = 0.0, s_d = 0.0;
    for (long long i = a; i < b; ++i) {
      const float n = float(c[i] + off);
      const double td = double(lp[i]) + double(beta * n), tn = double(logf(n)) + td;
      if (mx_n > -INFINITY) s_n += exp(tn - mx_n);
      s_d += exp(td - mx_d);
    }
    const double lse_n = (mx_n > -INFINITY) ? mx_n + log(s_n) : -INFINITY;
    const double outcome = lse_n - (mx_d + log(s_d)) - target;  // may be inf / nan exactly as in torch
    if (fabs(0.0 - outcome) < tol) break;
    if (outcome < -tol) lo = beta;
    if (outcome > tol) hi = beta;
  }
  if (beta_out != nullptr) beta_out[sidx] = beta;
  double mx = -INFINITY;
  for (long long i = a; i < b; ++i) mx = fmax(mx, double(lp[i]) + double(beta * float(c[i] + off)));
  double z = 0.0;
  for (long long i = a; i < b; ++i) z += exp(double(lp[i]) + double(beta * float(c[i] + off)) - mx);
  const double lse = mx + log(z);
  for (long long i = a; i < b; ++i) {
    const double v = double(lp[i]) + double(beta * float(c[i] + off)) - lse;
    lp_out[i] = v;
    keep[i] = (exp(v) != 0.0) ? 1 : 0;
  }
}

// Exponential tilt of every m-segment with a GIVEN factor (PRmu._chunked_compute_regularized_posterior, reference
// posterior.py:1305-1393): out_c = lp_c + beta n_c - LSE_c(lp_c + beta n_c), n_c = c + offset_m (fp32 product, fp64
// sums), beta = beta[0] (one overall factor) or beta[gene of m] (per-gene mode); keep = exp(out) != 0 in fp64.
__global__ void __launch_bounds__(256) posterior_tilt_kernel(
    const long long* __restrict__ m, const int* __restrict__ c, const float* __restrict__ lp,
    const long long* __restrict__ seg_start, long long n_seg, long long n_entries, const long long* __restrict__ off_m,
    const int* __restrict__ off_v, long long n_off, const float* __restrict__ beta, long long n_beta,
    long long total_n_genes, double* __restrict__ lp_out, int* __restrict__ keep) {
  const long long sidx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (sidx >= n_seg) return;
  const long long a = seg_start[sidx];
  const long long b = (sidx + 1 < n_seg) ? seg_start[sidx + 1] : n_entries;
  const long long key = m[a];
  const int off = lookup_offset(off_m, off_v, n_off, key);
  const float bt = (n_beta == 1) ? beta[0] : beta[key % total_n_genes];
  double mx = -INFINITY;
  for (long long i = a; i < b; ++i) mx = fmax(mx, double(lp[i]) + double(bt * float(c[i] + off)));
  double z = 0.0;
  for (long long i = a; i < b; ++i) z += exp(double(lp[i]) + double(bt * float(c[i] + off)) - mx);
  const double lse = mx + log(z);
  for (long long i = a; i < b; ++i) {
    const double v = double(lp[i]) + double(bt * float(c[i] + off)) - lse;
    lp_out[i] = v;
    keep[i] = (exp(v) != 0.0) ? 1 : 0;
  }
}

// CSR row pointer of the [n_rows, n_cols] output: indptr[r] = lower_bound(seg_m, r * n_cols)
__global__ void __launch_bounds__(256) csr_indptr_kernel(const long long* __restrict__ seg_m, long long n_seg,
                                                         long long n_rows, long long n_cols,
                                                         long long* __restrict__ indptr, int* __restrict__ col) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i <= n_rows) {
    const long long key = i * n_cols;
    long long lo = 0, hi = n_seg;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (seg_m[mid] < key) lo = mid + 1; else hi = mid;
    }
    indptr[i] = lo;
  }
  if (i < n_seg) col[i] = int(seg_m[i] % n_cols);
}

// ---------------------------------------------------------------------------------------------------------
// MCKP
// ---------------------------------------------------------------------------------------------------------
// per-gene sum of the MAP estimates (integer atomics: order-independent)
__global__ void __launch_bounds__(256) mckp_gene_sum_kernel(const long long* __restrict__ seg_m, const int* __restrict__ map_val,
                                                            long long n_seg, long long n_genes,
                                                            long long* __restrict__ gene_sum) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  atomicAdd(reinterpret_cast<unsigned long long*>(&gene_sum[seg_m[s] % n_genes]), (unsigned long long)(long long)map_val[s]);
}

// deficit[g] = (int64)(target[g] - gene_sum[g])   (numpy .astype(int): truncation toward zero)
__global__ void __launch_bounds__(256) mckp_deficit_kernel(const double* __restrict__ target, const long long* __restrict__ gene_sum,
                                                           long long n_genes, long long* __restrict__ deficit) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  deficit[g] = (long long)(target[g] - double(gene_sum[g]));
}

// candidate steps: key = (gene << 32) | float_bits(delta) for entries on the needed side of the MAP, else cand = 0.
// delta = |logp_i - logp_neighbour| in fp32 (estimation.py:633-643); non-finite deltas dropped (:655).
__global__ void __launch_bounds__(256) mckp_candidates_kernel(
    const long long* __restrict__ m, const int* __restrict__ c, const float* __restrict__ lp, const int* __restrict__ seg_of,
    const long long* __restrict__ map_pos, const long long* __restrict__ deficit, long long n_entries, long long n_genes,
    int* __restrict__ is_cand, unsigned long long* __restrict__ key) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const long long g = m[i] % n_genes;
  const long long d = deficit[g];
  const long long mp = map_pos[seg_of[i]];
  float delta = NAN;
  if (d > 0 && i > mp) delta = fabsf(lp[i] - lp[i - 1]);       // c > MAP: previous present entry of the same m
  else if (d < 0 && i < mp) delta = fabsf(lp[i] - lp[i + 1]);  // c < MAP: next present entry of the same m
  const bool ok = isfinite(delta);
  is_cand[i] = ok ? 1 : 0;
  key[i] = ok ? ((unsigned long long)g << 32) | (unsigned long long)__float_as_uint(delta) : ~0ULL;
}

// compact candidates (keys + original entry index) in entry order
__global__ void __launch_bounds__(256) mckp_compact_kernel(const int* __restrict__ is_cand, const long long* __restrict__ pos,
                                                           const unsigned long long* __restrict__ key, long long n_entries,
                                                           unsigned long long* __restrict__ cand_key,
                                                           long long* __restrict__ cand_idx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_entries || !is_cand[i]) return;
  cand_key[pos[i]] = key[i];
  cand_idx[pos[i]] = i;
}

// after a STABLE sort of candidates by key: rank inside the gene group < |deficit| -> one step for that entry's segment
__global__ void __launch_bounds__(256) mckp_select_kernel(const unsigned long long* __restrict__ sorted_key,
                                                          const long long* __restrict__ sorted_idx, long long n_cand,
                                                          const long long* __restrict__ deficit, const int* __restrict__ seg_of,
                                                          int* __restrict__ steps) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  const unsigned long long g = sorted_key[i] >> 32;
  // first position of this gene: lower_bound on (g << 32)
  const unsigned long long first_key = g << 32;
  long long lo = 0, hi = i;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (sorted_key[mid] < first_key) lo = mid + 1; else hi = mid;
  }
  const long long rank = i - lo;
  long long d = deficit[g];
  if (d < 0) d = -d;
  if (rank < d) atomicAdd(&steps[seg_of[sorted_idx[i]]], 1);
}

__global__ void __launch_bounds__(256) mckp_finalize_kernel(const long long* __restrict__ seg_m, const int* __restrict__ map_val,
                                                            const int* __restrict__ steps, const long long* __restrict__ deficit,
                                                            long long n_seg, long long n_genes, int* __restrict__ out) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const long long d = deficit[seg_m[s] % n_genes];
  const int dir = (d > 0) - (d < 0);
  out[s] = map_val[s] + dir * steps[s];
}

static inline unsigned blocks_for(long long n, int t = 256) { return unsigned((n + t - 1) / t); }

}  // namespace cb

// ---- C ABI ------------------------------------------------------------------------------------------------
// scratch: (n / 2048 + 2) int64
extern "C" int cb_exclusive_scan_i32(const int* in, long long n, long long* out, long long* scratch, long long* total_out,
                                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_tiles = (n + cb::kScanTile - 1) / cb::kScanTile;
  if (n == 0) {
    if (total_out) cudaMemsetAsync(total_out, 0, sizeof(long long), st);
    return 0;
  }
  cb::scan_tile_sums_kernel<<<unsigned(n_tiles), cb::kScanThreads, 0, st>>>(in, n, scratch);
  cb::scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(scratch, n_tiles);
  cb::scan_apply_kernel<<<unsigned(n_tiles), cb::kScanThreads, 0, st>>>(in, n, scratch, out);
  if (total_out) cudaMemcpyAsync(total_out, scratch + n_tiles, sizeof(long long), cudaMemcpyDeviceToDevice, st);
  return cb::check_launch("cb_exclusive_scan_i32");
}

extern "C" long long cb_scan_scratch_elems(long long n) { return n / cb::kScanTile + 3; }

extern "C" int cb_segment_heads(const long long* m, const int* c, long long n, int* head, int* unsorted_flag, void* stream) {
  if (n == 0) return 0;
  cb::segment_heads_kernel<<<cb::blocks_for(n), 256, 0, (cudaStream_t)stream>>>(m, c, n, head, unsorted_flag);
  return cb::check_launch("cb_segment_heads");
}

extern "C" int cb_segment_starts(const int* head, const long long* pos, long long n, long long* seg_start, int* seg_of,
                                 void* stream) {
  if (n == 0) return 0;
  cb::segment_starts_kernel<<<cb::blocks_for(n), 256, 0, (cudaStream_t)stream>>>(head, pos, n, seg_start, seg_of);
  return cb::check_launch("cb_segment_starts");
}

extern "C" int cb_segment_estimate(int mode, const long long* m, const int* c, const float* lp, const long long* seg_start,
                                   long long n_seg, long long n_entries, const long long* off_m, const int* off_v,
                                   long long n_off, double q, int n_cols, unsigned long long seed, long long* seg_m,
                                   int* out_i, float* out_f, long long* map_pos, void* stream) {
  if (n_seg == 0) return 0;
  CB_REQUIRE(mode >= 0 && mode <= 3, "cb_segment_estimate: unknown mode %d", mode);
  cb::segment_estimate_kernel<<<cb::blocks_for(n_seg), 256, 0, (cudaStream_t)stream>>>(
      mode, m, c, lp, seg_start, n_seg, n_entries, off_m, off_v, n_off, q, n_cols, seed, seg_m, out_i, out_f, map_pos);
  return cb::check_launch("cb_segment_estimate");
}

extern "C" int cb_prq_regularize(const long long* m, const int* c, const float* lp, const long long* seg_start,
                                 long long n_seg, long long n_entries, const long long* off_m, const int* off_v,
                                 long long n_off, double alpha, double target_tolerance, int max_iterations,
                                 double* log_target, float* beta_out, double* lp_out, int* keep, void* stream) {
  if (n_seg == 0) return 0;
  CB_REQUIRE(target_tolerance > 0.0 && max_iterations > 0, "cb_prq_regularize: target_tolerance should be > 0.");
  cb::prq_regularize_kernel<<<cb::blocks_for(n_seg), 256, 0, (cudaStream_t)stream>>>(
      m, c, lp, seg_start, n_seg, n_entries, off_m, off_v, n_off, alpha, target_tolerance, max_iterations, log_target,
      beta_out, lp_out, keep);
  return cb::check_launch("cb_prq_regularize");
}

extern "C" int cb_posterior_tilt(const long long* m, const int* c, const float* lp, const long long* seg_start,
                                 long long n_seg, long long n_entries, const long long* off_m, const int* off_v,
                                 long long n_off, const float* beta, long long n_beta, long long total_n_genes,
                                 double* lp_out, int* keep, void* stream) {
  if (n_seg == 0) return 0;
  CB_REQUIRE(n_beta == 1 || n_beta == total_n_genes, "cb_posterior_tilt: beta must hold 1 or total_n_genes (%lld) values, got %lld",
             total_n_genes, n_beta);
  cb::posterior_tilt_kernel<<<cb::blocks_for(n_seg), 256, 0, (cudaStream_t)stream>>>(
      m, c, lp, seg_start, n_seg, n_entries, off_m, off_v, n_off, beta, n_beta, total_n_genes, lp_out, keep);
  return cb::check_launch("cb_posterior_tilt");
}

extern "C" int cb_csr_from_segments(const long long* seg_m, long long n_seg, long long n_rows, long long n_cols,
                                    long long* indptr, int* col, void* stream) {
  const long long n = (n_rows + 1 > n_seg) ? n_rows + 1 : n_seg;
  cb::csr_indptr_kernel<<<cb::blocks_for(n), 256, 0, (cudaStream_t)stream>>>(seg_m, n_seg, n_rows, n_cols, indptr, col);
  return cb::check_launch("cb_csr_from_segments");
}

extern "C" int cb_mckp_gene_deficit(const long long* seg_m, const int* map_val, long long n_seg, const double* target,
                                    long long n_genes, long long* gene_sum, long long* deficit, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(gene_sum, 0, sizeof(long long) * n_genes, st);
  if (n_seg > 0) cb::mckp_gene_sum_kernel<<<cb::blocks_for(n_seg), 256, 0, st>>>(seg_m, map_val, n_seg, n_genes, gene_sum);
  cb::mckp_deficit_kernel<<<cb::blocks_for(n_genes), 256, 0, st>>>(target, gene_sum, n_genes, deficit);
  return cb::check_launch("cb_mckp_gene_deficit");
}

extern "C" int cb_mckp_candidates(const long long* m, const int* c, const float* lp, const int* seg_of,
                                  const long long* map_pos, const long long* deficit, long long n_entries,
                                  long long n_genes, int* is_cand, unsigned long long* key, void* stream) {
  if (n_entries == 0) return 0;
  cb::mckp_candidates_kernel<<<cb::blocks_for(n_entries), 256, 0, (cudaStream_t)stream>>>(m, c, lp, seg_of, map_pos, deficit,
                                                                                         n_entries, n_genes, is_cand, key);
  return cb::check_launch("cb_mckp_candidates");
}

extern "C" int cb_mckp_compact(const int* is_cand, const long long* pos, const unsigned long long* key, long long n_entries,
                               unsigned long long* cand_key, long long* cand_idx, void* stream) {
  if (n_entries == 0) return 0;
  cb::mckp_compact_kernel<<<cb::blocks_for(n_entries), 256, 0, (cudaStream_t)stream>>>(is_cand, pos, key, n_entries, cand_key,
                                                                                      cand_idx);
  return cb::check_launch("cb_mckp_compact");
}

extern "C" int cb_mckp_select(const unsigned long long* sorted_key, const long long* sorted_idx, long long n_cand,
                              const long long* deficit, const int* seg_of, int* steps, void* stream) {
  if (n_cand == 0) return 0;
  cb::mckp_select_kernel<<<cb::blocks_for(n_cand), 256, 0, (cudaStream_t)stream>>>(sorted_key, sorted_idx, n_cand, deficit,
                                                                                  seg_of, steps);
  return cb::check_launch("cb_mckp_select");
}

extern "C" int cb_mckp_finalize(const long long* seg_m, const int* map_val, const int* steps, const long long* deficit,
                                long long n_seg, long long n_genes, int* out, void* stream) {
  if (n_seg == 0) return 0;
  cb::mckp_finalize_kernel<<<cb::blocks_for(n_seg), 256, 0, (cudaStream_t)stream>>>(seg_m, map_val, steps, deficit, n_seg,
                                                                                   n_genes, out);
  return cb::check_launch("cb_mckp_finalize");
}
