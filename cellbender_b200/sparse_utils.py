"""Host mirror of the sparse-matrix helpers the reference uses to assemble the output count matrix
(cellbender/remove_background/sparse_utils.py:58-100, posterior.py:315-328, data/dataset.py:495-528), running on the
device through csrc/csr_combine.cu.  Same names and argument meaning as the reference; scipy matrices in, scipy
matrices out.  No CPU fallback: a missing CUDA extension or device raises.

All of them are instances of one row-wise combination

    out[n, g] = row_a[n] * col_a[g] * A[n, g] + sign * row_b[n] * col_b[g] * B[n, g]      (explicit zeros dropped)

evaluated by ``combine_csr`` in two kernel passes plus two prefix scans (integer / byte work, bit-exact).
"""
from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi, ops
from ._cabi import current_stream, ptr

_DTYPE_CODES = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
_TORCH_OF = {0: torch.int32, 1: torch.int64, 2: torch.float32, 3: torch.float64}


def _code_of(x: torch.Tensor) -> int:
    """dtype code of the C ABI (0 int32, 1 int64, 2 float32, 3 float64) of a tensor."""
    for code, dt in _TORCH_OF.items():
        if x.dtype == dt:
            return code
    raise TypeError(f"unsupported value dtype {x.dtype} (int32, int64, float32, float64)")


def _value_dtype(*dtypes) -> np.dtype:
    """numpy's result type of the operands, widened to one of the four value types the kernels are built for."""
    dt = np.result_type(*dtypes)
    if dt in _DTYPE_CODES:
        return dt
    if dt.kind in "bui":
        return np.dtype(np.int32) if np.can_cast(dt, np.int32) else np.dtype(np.int64)
    if dt.kind == "f":
        return np.dtype(np.float32) if dt.itemsize <= 4 else np.dtype(np.float64)
    raise TypeError(f"unsupported sparse value dtype {dt}")


@dataclass
class DeviceSparse:
    """Canonical CSR on the device: ptr int64[R+1], idx int32[nnz] sorted per row, val in one of the kernel dtypes."""

    ptr: torch.Tensor
    idx: torch.Tensor
    val: torch.Tensor
    shape: Tuple[int, int]

    @property
    def nnz(self) -> int:
        return int(self.idx.numel())

    @staticmethod
    def from_scipy(mat, dtype: np.dtype, device) -> "DeviceSparse":
        """Upload a scipy matrix.  A csc_matrix is uploaded as it is stored (the CSR form of its transpose) and
        transposed on the device, so no host-side format conversion happens for either layout."""
        transposed = isinstance(mat, sp.csc_matrix)
        if not isinstance(mat, (sp.csr_matrix, sp.csc_matrix)):
            mat = sp.csr_matrix(mat)
        if not mat.has_canonical_format:
            mat = mat.copy()
            mat.sum_duplicates()
        if mat.shape[0] >= 2 ** 31 or mat.shape[1] >= 2 ** 31:
            raise ValueError("matrix dimensions must fit int32 indices")
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device, non_blocking=False)
        shape = (int(mat.shape[1]), int(mat.shape[0])) if transposed else (int(mat.shape[0]), int(mat.shape[1]))
        out = DeviceSparse(up(mat.indptr, np.int64), up(mat.indices, np.int32), up(mat.data, dtype), shape)
        return out.transpose() if transposed else out

    @staticmethod
    def empty(shape, dtype: np.dtype, device) -> "DeviceSparse":
        code = _DTYPE_CODES[np.dtype(dtype)]
        return DeviceSparse(torch.zeros(shape[0] + 1, dtype=torch.int64, device=device),
                            torch.empty(0, dtype=torch.int32, device=device),
                            torch.empty(0, dtype=_TORCH_OF[code], device=device), (int(shape[0]), int(shape[1])))

    def transpose(self) -> "DeviceSparse":
        """CSR of the transpose (= CSC of this matrix) through cb_csr_to_csc; the stable sort of the column indices is
        torch's radix sort.  Canonical in, canonical out."""
        dev = self.idx.device
        n_rows, n_cols = self.shape
        nnz = self.nnz
        code = _code_of(self.val)
        sorted_col, order = torch.sort(self.idx, stable=True)
        col_ptr = torch.empty(n_cols + 1, dtype=torch.int64, device=dev)
        csc_row = torch.empty(nnz, dtype=torch.int32, device=dev)
        csc_val = torch.empty_like(self.val)
        _cabi.lib().call("cb_csr_to_csc", n_rows, n_cols, nnz, ptr(self.ptr), ptr(self.val), code, ptr(order),
                         ptr(sorted_col), ptr(col_ptr), ptr(csc_row), ptr(csc_val), current_stream())
        return DeviceSparse(col_ptr, csc_row, csc_val, (n_cols, n_rows))

    def _download(self, cls, shape):
        out = cls((self.val.cpu().numpy(), self.idx.cpu().numpy(), self.ptr.cpu().numpy()), shape=shape)
        out.has_canonical_format = True  # sorted, duplicate-free by construction
        return out

    def to_scipy_csr(self) -> sp.csr_matrix:
        return self._download(sp.csr_matrix, self.shape)

    def to_scipy_csc(self) -> sp.csc_matrix:
        """`.tocsc()` on the device."""
        return self.transpose()._download(sp.csc_matrix, self.shape)


def _mask(n: int, inds, device, invert: bool = False) -> torch.Tensor:
    m = np.zeros(n, dtype=np.uint8)
    inds = np.asarray(list(inds) if not isinstance(inds, np.ndarray) else inds, dtype=np.int64)
    if inds.size:
        m[inds] = 1
    if invert:
        m = 1 - m
    return torch.from_numpy(m).to(device)


def combine_device(a: DeviceSparse, b: Optional[DeviceSparse], sign: int = 1, row_a: Optional[torch.Tensor] = None,
                   row_b: Optional[torch.Tensor] = None, col_a: Optional[torch.Tensor] = None,
                   col_b: Optional[torch.Tensor] = None) -> DeviceSparse:
    """out = row_a * col_a * A + sign * row_b * col_b * B on device-resident canonical CSR operands."""
    ops._req_cuda(a.ptr, a.idx, a.val, row_a, row_b, col_a, col_b)
    dev = a.idx.device
    if b is None:
        b = DeviceSparse(torch.zeros_like(a.ptr), a.idx[:0], a.val[:0], a.shape)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.val.dtype != b.val.dtype:
        raise TypeError("operands must share one value dtype")
    for m, n in ((row_a, a.shape[0]), (row_b, a.shape[0]), (col_a, a.shape[1]), (col_b, a.shape[1])):
        if m is not None and (m.dtype != torch.uint8 or m.numel() != n):
            raise ValueError("masks are uint8 vectors over rows / columns")
    code = _code_of(a.val)
    L, st = _cabi.lib(), current_stream()
    n_rows = a.shape[0]
    flag_a = torch.empty(a.nnz + 1, dtype=torch.int32, device=dev)
    flag_b = torch.empty(b.nnz + 1, dtype=torch.int32, device=dev)
    comb_a = torch.empty_like(a.val)
    L.call("cb_csr_combine_flags", n_rows, ptr(a.ptr), ptr(a.idx), ptr(a.val), a.nnz, ptr(b.ptr), ptr(b.idx), ptr(b.val),
           b.nnz, ptr(row_a), ptr(row_b), ptr(col_a), ptr(col_b), int(sign), code, ptr(flag_a), ptr(flag_b), ptr(comb_a), st)
    scan_a, _ = ops.exclusive_scan_i32(flag_a)
    scan_b, _ = ops.exclusive_scan_i32(flag_b)
    nnz_out = int((scan_a[-1] + scan_b[-1]).item())
    out = DeviceSparse(torch.empty(n_rows + 1, dtype=torch.int64, device=dev),
                       torch.empty(nnz_out, dtype=torch.int32, device=dev),
                       torch.empty(nnz_out, dtype=a.val.dtype, device=dev), a.shape)
    L.call("cb_csr_combine_fill", n_rows, ptr(a.ptr), ptr(a.idx), ptr(b.ptr), ptr(b.idx), ptr(b.val), int(sign), code,
           ptr(flag_a), ptr(flag_b), ptr(scan_a), ptr(scan_b), ptr(comb_a), ptr(out.ptr), ptr(out.idx), ptr(out.val), st)
    return out


def combine_csr(mat_a, mat_b=None, sign: int = 1, keep_rows: Optional[Iterable[int]] = None,
                zero_rows: Optional[Iterable[int]] = None, rows_apply_to_b: bool = True,
                cols_a: Optional[Iterable[int]] = None, cols_b_complement: bool = False, device: str = "cuda",
                fmt: str = "csr"):
    """scipy in, scipy out.  ``keep_rows`` / ``zero_rows``: rows kept / rows emptied (at most one of them), in A and,
    if ``rows_apply_to_b``, in B; ``cols_a``: columns taken from A (all when None); ``cols_b_complement``: B contributes
    only the other columns."""
    if not torch.cuda.is_available():
        raise RuntimeError("cellbender_b200.sparse_utils needs a CUDA device (no CPU fallback)")
    dt = _value_dtype(mat_a.dtype, *(() if mat_b is None else (mat_b.dtype,)))
    a = DeviceSparse.from_scipy(mat_a, dt, device)
    b = None if mat_b is None else DeviceSparse.from_scipy(mat_b, dt, device)
    row_keep = None
    if keep_rows is not None:
        row_keep = _mask(a.shape[0], keep_rows, device)
    elif zero_rows is not None:
        row_keep = _mask(a.shape[0], zero_rows, device, invert=True)
    col_a = None if cols_a is None else _mask(a.shape[1], cols_a, device)
    col_b = (1 - col_a) if (cols_b_complement and col_a is not None) else None
    out = combine_device(a, b, sign, row_keep, row_keep if rows_apply_to_b else None, col_a, col_b)
    return out.to_scipy_csc() if fmt == "csc" else out.to_scipy_csr()


# ---- the reference's functions --------------------------------------------------------------------------------
@torch.no_grad()
def dense_to_sparse_op_torch(t, tensor_for_nonzeros: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, ...]:
    """Dense matrix -> COO tuple ``(*nonzero_inds, nonzero_values)`` (reference sparse_utils.py:10-33): indices int64 in
    torch.nonzero's row-major order; with ``tensor_for_nonzeros`` the positions come from that tensor and the values
    from ``t``.  Accepts a ``dataprep.MiniBatch`` (what this package's DataLoader yields) in place of a dense tensor."""
    t = t.dense() if hasattr(t, "dense") and not isinstance(t, torch.Tensor) else t
    mask = t if tensor_for_nonzeros is None else tensor_for_nonzeros
    mask = mask.dense() if hasattr(mask, "dense") and not isinstance(mask, torch.Tensor) else mask
    ops._req_cuda(t, mask)
    if tuple(mask.shape) != tuple(t.shape):
        raise ValueError("tensor_for_nonzeros must have the shape of t")
    if mask.dtype == torch.bool:
        mask = mask.to(torch.int32)
    shape = tuple(t.shape)
    if len(shape) == 0:
        raise ValueError("dense_to_sparse_op_torch needs a tensor with at least one dimension")

    def as2d(x):  # rows x last dimension, unit stride inside a row (a padded leading dimension is fine)
        x = x.reshape(1, -1) if x.dim() < 2 else x
        if x.dim() > 2 or x.stride(-1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
            x = x.contiguous().reshape(-1, x.shape[-1])
        return x

    t2, m2 = as2d(t), as2d(mask)
    codes = [_code_of(t2), _code_of(m2)]
    n_rows, n_cols = int(t2.shape[0]), int(t2.shape[1])
    ld = lambda x: int(x.stride(0)) if x.shape[0] > 1 else max(n_cols, 1)
    dev = t.device
    L, st = _cabi.lib(), current_stream()
    flag = torch.empty(n_rows * n_cols, dtype=torch.int32, device=dev)
    L.call("cb_dense_nonzero_flags", ptr(m2), codes[1], n_rows, n_cols, ld(m2), ptr(flag), st)
    pos, total = ops.exclusive_scan_i32(flag)
    nnz = int(total.item())
    row = torch.empty(nnz, dtype=torch.int64, device=dev)
    col = torch.empty(nnz, dtype=torch.int64, device=dev)
    val = torch.empty(nnz, dtype=t.dtype, device=dev)
    L.call("cb_dense_to_coo", ptr(t2), codes[0], n_rows, n_cols, ld(t2), ptr(flag), ptr(pos), ptr(row), ptr(col), ptr(val), st)
    if len(shape) == 2:
        return row, col, val
    if len(shape) == 1:
        return col, val
    lead = torch.unravel_index(row, shape[:-1])  # N-d input: split the flattened leading index
    return tuple(lead) + (col, val)


def csr_set_rows_to_zero(csr, row_inds) -> sp.csr_matrix:
    """Set all nonzero elements in rows ``row_inds`` to zero (reference sparse_utils.py:58-73).  Happens in place
    for a csr_matrix input, like the reference; the result is returned as well."""
    if not isinstance(csr, sp.csr_matrix):
        if sp.issparse(csr):
            return combine_csr(csr, zero_rows=row_inds).astype(csr.dtype, copy=False)  # format conversion on the device
        try:
            csr = csr.tocsr()
        except Exception:
            raise ValueError("Matrix given must be of CSR format.")
    out = combine_csr(csr, zero_rows=row_inds)
    out = out.astype(csr.dtype, copy=False)
    csr.data, csr.indices, csr.indptr = out.data, out.indices.astype(csr.indices.dtype, copy=False), \
        out.indptr.astype(csr.indptr.dtype, copy=False)
    return csr


def overwrite_matrix_with_columns_from_another(mat1, mat2, column_inds) -> sp.csc_matrix:
    """Replace the columns of ``mat1`` that are NOT in ``column_inds`` with the entries of ``mat2``
    (reference sparse_utils.py:76-100)."""
    return combine_csr(mat1, mat2, sign=1, cols_a=column_inds, cols_b_complement=True, fmt="csc")


def subtract_noise_in_cells(count_matrix, noise_csr, cell_inds) -> sp.csc_matrix:
    """``csr_set_rows_to_zero(count_matrix, not cells) - noise_csr`` as CSC (reference posterior.py:315-328)."""
    return combine_csr(count_matrix, noise_csr, sign=-1, keep_rows=cell_inds, rows_apply_to_b=False, fmt="csc")


def restore_eliminated_features_in_cells(inferred_count_matrix, raw_count_matrix, analyzed_gene_inds,
                                         analyzed_barcode_inds, cell_probabilities_analyzed_bcs,
                                         cell_prob_cutoff: float = 0.5) -> sp.csc_matrix:
    """SingleCellRNACountsDataset.restore_eliminated_features_in_cells (reference data/dataset.py:495-528): features
    not used during inference come back from the raw data; droplets with p <= cutoff (and all droplets that were never
    analysed) are empty in the output.  One fused pass instead of overwrite + csr_set_rows_to_zero."""
    p_all = np.zeros(raw_count_matrix.shape[0])
    p_all[np.asarray(analyzed_barcode_inds)] = np.asarray(cell_probabilities_analyzed_bcs)
    return combine_csr(inferred_count_matrix, raw_count_matrix, sign=1, keep_rows=np.flatnonzero(p_all > cell_prob_cutoff),
                       cols_a=analyzed_gene_inds, cols_b_complement=True, fmt="csc")
