// Row-wise combination of two CSR count matrices of equal shape, the output-assembly step after the estimators:
//
//     out[n, g] = row_a[n] * col_a[g] * A[n, g]  +  sign * row_b[n] * col_b[g] * B[n, g],      explicit zeros dropped,
//
// which covers, in the reference,
//   * csr_set_rows_to_zero                       (sparse_utils.py:58-73)     B empty, row_a = not in row_inds
//   * overwrite_matrix_with_columns_from_another (sparse_utils.py:76-100)    col_a = in column_inds, col_b = its complement
//   * cell_counts - noise_csr                    (posterior.py:321-326)      sign = -1, row_a = called cells (B unmasked)
//   * restore_eliminated_features_in_cells       (data/dataset.py:495-528)   the overwrite with row_a = row_b = p > 0.5
// The reference does these with Python loops over every stored index (`[i in set for i in mat.indices]`) and scipy
// binops on the host.  Integer / byte work, HBM-bound: per stored entry one read of (index, value), one 4-byte flag,
// one 8-byte scan value and one write of (index, value); no atomics, so the output is deterministic and canonical
// (sorted indices, no duplicates, no explicit zeros) exactly like scipy's `A + B` followed by eliminate_zeros().
//
// Two passes, one warp per row.  Pass 1 decides for every stored entry whether it survives into the output (A entries
// absorb the B entry of the same column; a B entry survives on its own only where A stores nothing usable) and keeps
// the combined value.  After exclusive scans SA, SB of the two flag arrays (cb_exclusive_scan_i32), the output position
// of a surviving A entry i with column g is SA[i] + SB[lower_bound_B(g)], of a surviving B entry j it is
// SB[j] + SA[lower_bound_A(g)], and the output row pointer is SA[ptrA[n]] + SB[ptrB[n]].
#include "common.cuh"

namespace cb {

constexpr int kCombineThreads = 256;
constexpr int kCombineRowsPerBlock = kCombineThreads / 32;

// first index in [lo, hi) whose column is >= g
__device__ __forceinline__ long long lower_bound_col(const int* __restrict__ idx, long long lo, long long hi, int g) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(idx + mid) < g) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(kCombineThreads)
csr_combine_flag_kernel(long long n_rows, const long long* __restrict__ ptr_a, const int* __restrict__ idx_a,
                        const T* __restrict__ val_a, const long long* __restrict__ ptr_b, const int* __restrict__ idx_b,
                        const T* __restrict__ val_b, const unsigned char* __restrict__ row_a,
                        const unsigned char* __restrict__ row_b, const unsigned char* __restrict__ col_a, const unsigned char* __restrict__ col_b, int sign,
                        int* __restrict__ flag_a, int* __restrict__ flag_b, T* __restrict__ comb_a) {
  const long long row = (long long)blockIdx.x * kCombineRowsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const long long a0 = ptr_a[row], a1 = ptr_a[row + 1], b0 = ptr_b[row], b1 = ptr_b[row + 1];
  const bool rk_a = row_a == nullptr || row_a[row] != 0;
  const bool rk_b = row_b == nullptr || row_b[row] != 0;
  for (long long i = a0 + lane; i < a1; i += 32) {
    const int g = idx_a[i];
    T v = T(0);
    bool keep = rk_a && (col_a == nullptr || col_a[g] != 0);
    if (keep) {
      v = val_a[i];
      if (rk_b && (col_b == nullptr || col_b[g] != 0)) {
        const long long j = lower_bound_col(idx_b, b0, b1, g);
        if (j < b1 && idx_b[j] == g) v = sign > 0 ? T(v + val_b[j]) : T(v - val_b[j]);
      }
      keep = v != T(0);
    }
    flag_a[i] = keep ? 1 : 0;
    comb_a[i] = v;
  }
  for (long long j = b0 + lane; j < b1; j += 32) {
    const int g = idx_b[j];
    bool keep = rk_b && (col_b == nullptr || col_b[g] != 0) && val_b[j] != T(0);
    if (keep && rk_a && (col_a == nullptr || col_a[g] != 0)) {  // absorbed by A's entry of the same column, if A stores one
      const long long i = lower_bound_col(idx_a, a0, a1, g);
      if (i < a1 && idx_a[i] == g) keep = false;
    }
    flag_b[j] = keep ? 1 : 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(kCombineThreads)
csr_combine_fill_kernel(long long n_rows, const long long* __restrict__ ptr_a, const int* __restrict__ idx_a,
                        const long long* __restrict__ ptr_b, const int* __restrict__ idx_b, const T* __restrict__ val_b,
                        int sign, const int* __restrict__ flag_a, const int* __restrict__ flag_b,
                        const long long* __restrict__ scan_a, const long long* __restrict__ scan_b,
                        const T* __restrict__ comb_a, long long* __restrict__ out_ptr, int* __restrict__ out_idx,
                        T* __restrict__ out_val) {
  const long long row = (long long)blockIdx.x * kCombineRowsPerBlock + (threadIdx.x >> 5);
  if (row > n_rows) return;  // row == n_rows writes the closing row pointer
  const int lane = threadIdx.x & 31;
  const long long a0 = ptr_a[row], b0 = ptr_b[row];
  if (lane == 0) out_ptr[row] = scan_a[a0] + scan_b[b0];
  if (row == n_rows) return;
  const long long a1 = ptr_a[row + 1], b1 = ptr_b[row + 1];
  for (long long i = a0 + lane; i < a1; i += 32) {
    if (!flag_a[i]) continue;
    const int g = idx_a[i];
    const long long pos = scan_a[i] + scan_b[lower_bound_col(idx_b, b0, b1, g)];
    out_idx[pos] = g;
    out_val[pos] = comb_a[i];
  }
  for (long long j = b0 + lane; j < b1; j += 32) {
    if (!flag_b[j]) continue;
    const int g = idx_b[j];
    const long long pos = scan_b[j] + scan_a[lower_bound_col(idx_a, a0, a1, g)];
    out_idx[pos] = g;
    out_val[pos] = sign > 0 ? val_b[j] : T(-val_b[j]);
  }
}

// CSR -> CSC of a canonical matrix, given the stable order of the stored entries by column (`order`, from a stable
// sort of the column indices): row of every entry by binary search in the row pointer, values permuted.
template <typename T>
__global__ void csr_entries_to_csc_kernel(long long nnz, long long n_rows, const long long* __restrict__ ptr,
                                          const T* __restrict__ val, const long long* __restrict__ order,
                                          int* __restrict__ csc_row, T* __restrict__ csc_val) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const long long e = order[k];
  long long lo = 0, hi = n_rows;  // last row with ptr[row] <= e
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= e) lo = mid;
    else hi = mid;
  }
  csc_row[k] = int(lo);
  csc_val[k] = val[e];
}

// column pointer of the CSC form: col_ptr[g] = first position k with sorted_col[k] >= g
__global__ void csc_col_ptr_kernel(long long nnz, long long n_cols, const int* __restrict__ sorted_col,
                                   long long* __restrict__ col_ptr) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_cols) return;
  col_ptr[g] = g == n_cols ? nnz : lower_bound_col(sorted_col, 0, nnz, int(g));
}

// ---- dense -> COO (dense_to_sparse_op_torch, sparse_utils.py:10-33) ------------------------------------------------
// flag[r * n_cols + c] = (mask[r, c] != 0); after an exclusive scan of the flags the second kernel writes the
// (row, col, value) triplets in row-major order, the order torch.nonzero returns.
template <typename T>
__global__ void dense_nonzero_flag_kernel(const T* __restrict__ mask, long long n_rows, long long n_cols, long long ld,
                                          int* __restrict__ flag) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_rows * n_cols) return;
  const long long r = k / n_cols, c = k - r * n_cols;
  flag[k] = mask[r * ld + c] != T(0) ? 1 : 0;
}

template <typename T>
__global__ void dense_to_coo_kernel(const T* __restrict__ t, long long n_rows, long long n_cols, long long ld,
                                    const int* __restrict__ flag, const long long* __restrict__ pos,
                                    long long* __restrict__ row, long long* __restrict__ col, T* __restrict__ val) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_rows * n_cols || !flag[k]) return;
  const long long r = k / n_cols, c = k - r * n_cols, o = pos[k];
  row[o] = r;
  col[o] = c;
  val[o] = t[r * ld + c];
}

}  // namespace cb

#define CB_COMBINE_DISPATCH(dtype, ...)                    \
  switch (dtype) {                                         \
    case 0: { using T = int; __VA_ARGS__; } break;         \
    case 1: { using T = long long; __VA_ARGS__; } break;   \
    case 2: { using T = float; __VA_ARGS__; } break;       \
    case 3: { using T = double; __VA_ARGS__; } break;      \
    default: CB_REQUIRE(false, "csr combine: dtype code %d (0 int32, 1 int64, 2 float32, 3 float64)", dtype); \
  }

// flag_a[nnz_a + 1], flag_b[nnz_b + 1] (the extra trailing element is set to 0 so that the scans deliver the totals at
// index nnz); comb_a[nnz_a] holds the combined values of A's entries.
extern "C" int cb_csr_combine_flags(long long n_rows, const long long* ptr_a, const int* idx_a, const void* val_a,
                                    long long nnz_a, const long long* ptr_b, const int* idx_b, const void* val_b,
                                    long long nnz_b, const unsigned char* row_a, const unsigned char* row_b,
                                    const unsigned char* col_a, const unsigned char* col_b, int sign, int dtype, int* flag_a, int* flag_b, void* comb_a,
                                    void* stream) {
  CB_REQUIRE(n_rows >= 0 && nnz_a >= 0 && nnz_b >= 0 && (sign == 1 || sign == -1), "cb_csr_combine_flags: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(flag_a + nnz_a, 0, sizeof(int), st);
  cudaMemsetAsync(flag_b + nnz_b, 0, sizeof(int), st);
  if (n_rows == 0) return 0;
  const unsigned grid = unsigned((n_rows + cb::kCombineRowsPerBlock - 1) / cb::kCombineRowsPerBlock);
  CB_COMBINE_DISPATCH(dtype, cb::csr_combine_flag_kernel<T><<<grid, cb::kCombineThreads, 0, st>>>(
                                 n_rows, ptr_a, idx_a, (const T*)val_a, ptr_b, idx_b, (const T*)val_b, row_a, row_b, col_a, col_b,
                                 sign, flag_a, flag_b, (T*)comb_a));
  return cb::check_launch("cb_csr_combine_flags");
}

// scan_a[nnz_a + 1], scan_b[nnz_b + 1]: exclusive scans of the flag arrays (including the trailing zero).
// out_ptr[n_rows + 1]; out_idx / out_val sized scan_a[nnz_a] + scan_b[nnz_b].
extern "C" int cb_csr_combine_fill(long long n_rows, const long long* ptr_a, const int* idx_a, const long long* ptr_b,
                                   const int* idx_b, const void* val_b, int sign, int dtype, const int* flag_a,
                                   const int* flag_b, const long long* scan_a, const long long* scan_b, const void* comb_a,
                                   long long* out_ptr, int* out_idx, void* out_val, void* stream) {
  CB_REQUIRE(n_rows >= 0 && (sign == 1 || sign == -1), "cb_csr_combine_fill: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = unsigned((n_rows + 1 + cb::kCombineRowsPerBlock - 1) / cb::kCombineRowsPerBlock);
  CB_COMBINE_DISPATCH(dtype, cb::csr_combine_fill_kernel<T><<<grid, cb::kCombineThreads, 0, st>>>(
                                 n_rows, ptr_a, idx_a, ptr_b, idx_b, (const T*)val_b, sign, flag_a, flag_b, scan_a, scan_b,
                                 (const T*)comb_a, out_ptr, out_idx, (T*)out_val));
  return cb::check_launch("cb_csr_combine_fill");
}

// CSR -> CSC (what `.tocsc()` does on the host at posterior.py:328 and dataset.py:528).  order[nnz]: permutation that
// stably sorts the entries by column; sorted_col[nnz] = idx[order].  Outputs col_ptr[n_cols + 1], csc_row[nnz], csc_val[nnz].
extern "C" int cb_csr_to_csc(long long n_rows, long long n_cols, long long nnz, const long long* ptr, const void* val,
                             int dtype, const long long* order, const int* sorted_col, long long* col_ptr, int* csc_row,
                             void* csc_val, void* stream) {
  CB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "cb_csr_to_csc: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cb::csc_col_ptr_kernel<<<unsigned((n_cols + 1 + 255) / 256), 256, 0, st>>>(nnz, n_cols, sorted_col, col_ptr);
  if (nnz > 0) {
    CB_COMBINE_DISPATCH(dtype, cb::csr_entries_to_csc_kernel<T><<<unsigned((nnz + 255) / 256), 256, 0, st>>>(
                                   nnz, n_rows, ptr, (const T*)val, order, csc_row, (T*)csc_val));
  }
  return cb::check_launch("cb_csr_to_csc");
}

// dense_to_sparse_op_torch (sparse_utils.py:10-33) in two launches around cb_exclusive_scan_i32: mask / t are row-major
// [n_rows, n_cols] with leading dimensions ld_mask / ld_t (elements); flag[n_rows * n_cols]; pos = exclusive scan of flag;
// row / col [nnz] int64 like torch.nonzero, val[nnz] of t's type.  dtype codes as above (mask and t separately).
extern "C" int cb_dense_nonzero_flags(const void* mask, int dtype, long long n_rows, long long n_cols, long long ld_mask,
                                      int* flag, void* stream) {
  CB_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld_mask >= n_cols, "cb_dense_nonzero_flags: bad shape");
  const long long n = n_rows * n_cols;
  if (n == 0) return 0;
  CB_COMBINE_DISPATCH(dtype, cb::dense_nonzero_flag_kernel<T><<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                                 (const T*)mask, n_rows, n_cols, ld_mask, flag));
  return cb::check_launch("cb_dense_nonzero_flags");
}

extern "C" int cb_dense_to_coo(const void* t, int dtype, long long n_rows, long long n_cols, long long ld_t, const int* flag,
                               const long long* pos, long long* row, long long* col, void* val, void* stream) {
  CB_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld_t >= n_cols, "cb_dense_to_coo: bad shape");
  const long long n = n_rows * n_cols;
  if (n == 0) return 0;
  CB_COMBINE_DISPATCH(dtype, cb::dense_to_coo_kernel<T><<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                                 (const T*)t, n_rows, n_cols, ld_t, flag, pos, row, col, (T*)val));
  return cb::check_launch("cb_dense_to_coo");
}
