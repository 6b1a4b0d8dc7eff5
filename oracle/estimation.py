"""ORACLE (test infrastructure, NOT the product path): numpy restatement of the reference's
noise-count estimators as segmented reductions over the (m, c)-sorted posterior COO.

Reference locations (relative to /root/reference/cellbender/remove_background):
  apply_function_dense_chunks / chunked_iterator   estimation.py:705-793
  MAP / Mean / ThresholdCDF / SingleSample         estimation.py:89-204
  _estimation_array_to_csr                         estimation.py:207-235
  MultipleChoiceKnapsack._chunk_estimate_noise     estimation.py:552-702
  log_prob_sparse_to_dense / todense_fill          sparse_utils.py:36-55
  compute_mean_target_removal_as_function          posterior.py:1589-1650

Semantics restated (see SURVEY.md section 8a "Verified restatements"):
  * every estimator densifies a segment (all entries of one m) through float64 with
    -inf fill; MAP = first argmax; CDF = #{c : cumsum_{c'<=c} exp(logp) <= q} over the
    dense columns 0..C-1 (C = number of distinct columns present = max c + 1 for
    realistic posteriors); Mean = sum_c c*exp(logp) in float64, + offset, cast float32.
  * the reference re-indexes columns compactly per chunk and ignores the mapping back
    (estimation.py:739,778); this restatement uses the true c, equal whenever every
    c in [0, max c] occurs in the COO (true for every reference test and real data).
  * MCKP: MAP -> per-gene deficit int(target - sum MAP) -> candidates on the needed side
    of the MAP, delta = |logp_i - logp_neighbour| in float32 -> per gene the |deficit|
    smallest deltas, ties by (m, c) order -> noise = MAP +/- count per m.

Pinned against the reference's own known-answer tests (tests/test_estimation.py:25-59,
283-343) via tests/golden/estimation_*.npz and, when /root/reference is present, against
the imported reference on random posteriors (tests/test_oracle_vs_reference.py).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.sparse as sp


def _sorted_segments(coo: sp.coo_matrix):
    """Sort entries by (m, c); return arrays + segment starts."""
    m = np.asarray(coo.row).astype(np.int64)
    c = np.asarray(coo.col).astype(np.int64)
    lp = np.asarray(coo.data)
    order = np.lexsort((c, m))
    m, c, lp = m[order], c[order], lp[order]
    if m.size == 0:
        return m, c, lp, np.zeros(0, np.int64)
    head = np.ones(m.size, bool)
    head[1:] = m[1:] != m[:-1]
    return m, c, lp, np.flatnonzero(head)


def _offsets_for(m_unique: np.ndarray, noise_offsets: Optional[Dict[int, int]]) -> np.ndarray:
    if noise_offsets is None:
        return np.zeros(m_unique.size, np.int64)
    return np.array([noise_offsets.get(int(i), 0) for i in m_unique], dtype=np.int64)


def _to_csr(values, m_unique, shape, dtype) -> sp.csr_matrix:
    """_estimation_array_to_csr (estimation.py:207-235): (n, g) = divmod(m, G)."""
    n, g = np.divmod(m_unique, shape[1])
    out = sp.coo_matrix((np.asarray(values).astype(dtype), (n.astype(np.int32), g.astype(np.int32))),
                        shape=shape, dtype=dtype)
    out.sum_duplicates()
    return out.tocsr()


def segment_map(coo) -> Tuple[np.ndarray, np.ndarray]:
    """Per-m first argmax (column value c) -- MAP.torch_argmax on the dense chunk."""
    m, c, lp, starts = _sorted_segments(coo)
    ends = np.append(starts[1:], m.size)
    out = np.zeros(starts.size, np.int64)
    lp64 = lp.astype(np.float64)
    for i, (a, b) in enumerate(zip(starts, ends)):
        out[i] = c[a + int(np.argmax(lp64[a:b]))]  # np.argmax returns the first maximum
    return m[starts], out


def estimate_map(coo, noise_offsets, shape) -> sp.csr_matrix:
    mu, val = segment_map(coo)
    return _to_csr(val + _offsets_for(mu, noise_offsets), mu, shape, np.int32)


def estimate_mean(coo, noise_offsets, shape) -> sp.csr_matrix:
    m, c, lp, starts = _sorted_segments(coo)
    ends = np.append(starts[1:], m.size)
    val = np.zeros(starts.size, np.float64)
    p = np.exp(lp.astype(np.float64))
    for i, (a, b) in enumerate(zip(starts, ends)):
        val[i] = float(np.sum(p[a:b] * c[a:b]))
    mu = m[starts]
    return _to_csr(val + _offsets_for(mu, noise_offsets), mu, shape, np.float32)


def estimate_cdf(coo, noise_offsets, shape, q: float = 0.5) -> sp.csr_matrix:
    m, c, lp, starts = _sorted_segments(coo)
    ends = np.append(starts[1:], m.size)
    n_cols = int(c.max()) + 1 if c.size else 0
    val = np.zeros(starts.size, np.int64)
    for i, (a, b) in enumerate(zip(starts, ends)):
        dense = np.zeros(n_cols, np.float64)
        dense[c[a:b]] = np.exp(lp[a:b].astype(np.float64))
        val[i] = int((np.cumsum(dense) <= q).sum())
    mu = m[starts]
    return _to_csr(val + _offsets_for(mu, noise_offsets), mu, shape, np.int32)


def estimate_mckp(coo, noise_offsets, shape, noise_targets_per_gene: np.ndarray) -> sp.csr_matrix:
    """MultipleChoiceKnapsack.estimate_noise (estimation.py:423-702) as a segmented top-k.
    Gene chunks are independent (estimation.py:531-550) so one pass over all genes is
    identical to the chunked sum."""
    G = shape[1]
    assert noise_targets_per_gene.size == G
    m, c, lp, starts = _sorted_segments(coo)
    lp = lp.astype(np.float32)
    ends = np.append(starts[1:], m.size)
    mu = m[starts]
    seg_of = np.repeat(np.arange(starts.size), ends - starts)
    # MAP per segment (first argmax), with offsets, summed per gene.
    kstar = np.zeros(starts.size, np.int64)
    for i, (a, b) in enumerate(zip(starts, ends)):
        kstar[i] = a + int(np.argmax(lp[a:b].astype(np.float64)))
    map_c = c[kstar]
    map_val = map_c + _offsets_for(mu, noise_offsets)
    map_csr = _to_csr(map_val, mu, shape, np.int32)
    map_per_gene = np.array(map_csr.sum(axis=0)).squeeze()
    deficit = (noise_targets_per_gene - map_per_gene).astype(int)  # truncation toward zero
    g_of_entry = (m % G).astype(np.int64)
    direction = np.sign(deficit)[g_of_entry]

    # delta to the neighbouring PRESENT entry of the same m, float32 arithmetic.
    idx = np.arange(m.size)
    delta = np.full(m.size, np.nan, np.float32)
    pos = (direction > 0) & (c > map_c[seg_of])
    neg = (direction < 0) & (c < map_c[seg_of])
    with np.errstate(invalid="ignore"):
        delta[pos] = np.abs(lp[idx[pos]] - lp[idx[pos] - 1])
        delta[neg] = np.abs(lp[idx[neg]] - lp[idx[neg] + 1])
    cand = np.flatnonzero(np.isfinite(delta))
    steps = np.zeros(starts.size, np.int64)
    if cand.size:
        # per gene: |deficit| smallest deltas, ties -> earlier (m, c) position (stable sort)
        order = cand[np.lexsort((cand, delta[cand], g_of_entry[cand]))]
        g_sorted = g_of_entry[order]
        first = np.ones(order.size, bool)
        first[1:] = g_sorted[1:] != g_sorted[:-1]
        grp_start = np.maximum.accumulate(np.where(first, np.arange(order.size), 0))
        rank = np.arange(order.size) - grp_start
        take = order[rank < np.abs(deficit)[g_sorted]]
        np.add.at(steps, seg_of[take], 1)
    seg_dir = np.sign(deficit)[(mu % G).astype(np.int64)]
    return _to_csr(map_val + seg_dir * steps, mu, shape, np.int32)


def mean_target_removal_per_gene(coo, noise_offsets, shape, raw_count_csr_for_cells: sp.csr_matrix,
                                 n_cells: int, fpr: float) -> np.ndarray:
    """compute_mean_target_removal_as_function(per_gene=True) evaluated at ``fpr``
    (posterior.py:1623-1648).  Returns the per-cell-scaled target, as the reference does."""
    mean_csr = estimate_mean(coo, noise_offsets, shape)
    approx_signal = raw_count_csr_for_cells - mean_csr
    target = np.array(mean_csr.sum(axis=0)).squeeze()
    target = target + fpr * np.array(approx_signal.sum(axis=0)).squeeze()
    return target / n_cells
