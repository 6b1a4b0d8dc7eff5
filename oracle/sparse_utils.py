"""ORACLE (test infrastructure, NOT the product path): numpy / scipy restatement of the reference's output-assembly
helpers, written as dense algebra so that it shares no structure with the device kernels it checks.

Reference locations (relative to /root/reference/cellbender/remove_background):
  csr_set_rows_to_zero                          sparse_utils.py:58-73
  overwrite_matrix_with_columns_from_another    sparse_utils.py:76-100
  cell_counts - noise_csr                       posterior.py:315-328
  restore_eliminated_features_in_cells          data/dataset.py:495-528

Pinned against the unmodified reference through tests/golden/sparse_utils.npz (tools/make_golden.py:golden_sparse_utils),
which includes the three cases of the reference's own tests/test_sparse_utils.py:128-139.
Dense arithmetic: only for the small / medium cases tests use.
"""
import numpy as np
import scipy.sparse as sp

CELL_PROB_CUTOFF = 0.5  # consts.py:  CELL_PROB_CUTOFF


def _dense(m) -> np.ndarray:
    return np.asarray(m.todense()) if sp.issparse(m) else np.asarray(m)


def csr_set_rows_to_zero(mat, row_inds) -> sp.csr_matrix:
    """sparse_utils.py:58-73: rows in row_inds become empty; explicit zeros are eliminated."""
    d = _dense(mat).copy()
    rows = np.asarray(sorted(set(int(r) for r in row_inds)), dtype=np.int64)
    if rows.size:
        d[rows, :] = 0
    return sp.csr_matrix(d)


def overwrite_matrix_with_columns_from_another(mat1, mat2, column_inds) -> sp.csc_matrix:
    """sparse_utils.py:76-100: columns in column_inds come from mat1, every other column from mat2."""
    d1, d2 = _dense(mat1), _dense(mat2)
    sel = np.zeros(d1.shape[1], dtype=bool)
    sel[np.asarray(column_inds, dtype=np.int64)] = True
    return sp.csc_matrix(np.where(sel[None, :], d1, d2))


def denoised_counts(count_matrix, noise_csr, cell_inds) -> sp.csc_matrix:
    """posterior.py:315-328: observed counts of called cells minus their noise; all other droplets empty."""
    d = _dense(count_matrix)
    keep = np.zeros(d.shape[0], dtype=bool)
    keep[np.asarray(cell_inds, dtype=np.int64)] = True
    return sp.csc_matrix(np.where(keep[:, None], d, 0) - _dense(noise_csr))


def restore_eliminated_features_in_cells(inferred, raw, analyzed_gene_inds, analyzed_barcode_inds, cell_probabilities):
    """data/dataset.py:495-528: features not analysed are restored from the raw data, then droplets with
    p <= CELL_PROB_CUTOFF (and droplets never analysed, p = 0) are emptied."""
    out = _dense(overwrite_matrix_with_columns_from_another(inferred, raw, analyzed_gene_inds)).copy()
    p_all = np.zeros(out.shape[0])
    p_all[np.asarray(analyzed_barcode_inds)] = np.asarray(cell_probabilities)
    out[p_all <= CELL_PROB_CUTOFF, :] = 0
    return sp.csc_matrix(out)
