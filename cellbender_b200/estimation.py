"""Noise-count estimators on the GPU, behind the reference's ``EstimationMethod`` interface
(cellbender/remove_background/estimation.py:29-86): same class names, same
``estimate_noise(noise_log_prob_coo, noise_offsets, **kwargs) -> scipy.sparse.csr_matrix`` signature, same
kwargs (``device``, ``q``, ``noise_targets_per_gene``, ``n_chunks``, ``use_multiple_processes``), same errors.

The segmented reductions run in ``libcellbender_b200.so`` (csrc/estimation.cu); torch supplies device memory
and, for COOs that arrive unsorted and for the MCKP per-gene selection, its radix sort (plumbing).  A
posterior that is already on the device (``DevicePosteriorCOO``) never visits the host.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from ._cabi import current_stream, ptr
from .ops import exclusive_scan_i32

N_CELLS_DATATYPE = np.int32  # estimation.py:24-26
N_GENES_DATATYPE = np.int32
COUNT_DATATYPE = np.int32

MODE_MAP, MODE_MEAN, MODE_CDF, MODE_SAMPLE = 0, 1, 2, 3


class IndexConverter:
    """(n, g) <-> flattened m = n * total_n_genes + g  (reference posterior.py:1533-1586)."""

    def __init__(self, total_n_cells: int, total_n_genes: int):
        self.total_n_cells = int(total_n_cells)
        self.total_n_genes = int(total_n_genes)
        self.matrix_shape = (self.total_n_cells, self.total_n_genes)

    def __repr__(self):
        return (f"IndexConverter with\n\ttotal_n_cells: {self.total_n_cells}\n\ttotal_n_genes: {self.total_n_genes}"
                f"\n\tmatrix_shape: {self.matrix_shape}")

    def get_m_indices(self, cell_inds, gene_inds):
        if not ((cell_inds >= 0) & (cell_inds < self.total_n_cells)).all():
            raise ValueError(f"Requested cell_inds out of range: "
                             f"{cell_inds[(cell_inds < 0) | (cell_inds >= self.total_n_cells)]}")
        if not ((gene_inds >= 0) & (gene_inds < self.total_n_genes)).all():
            raise ValueError(f"Requested gene_inds out of range: "
                             f"{gene_inds[(gene_inds < 0) | (gene_inds >= self.total_n_genes)]}")
        if isinstance(cell_inds, np.ndarray):
            return cell_inds.astype(np.uint64) * self.total_n_genes + gene_inds.astype(np.uint64)
        if isinstance(cell_inds, torch.Tensor):
            return cell_inds.type(torch.int64) * self.total_n_genes + gene_inds.type(torch.int64)
        raise ValueError("IndexConverter.get_m_indices received cell_inds of unkown object type")

    def get_ng_indices(self, m_inds: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        if not ((m_inds >= 0) & (m_inds < self.total_n_cells * self.total_n_genes)).all():
            raise ValueError(f"Requested m_inds out of range: "
                             f"{m_inds[(m_inds < 0) | (m_inds >= self.total_n_cells * self.total_n_genes)]}")
        return np.divmod(m_inds, self.total_n_genes)


@dataclass
class DevicePosteriorCOO:
    """The posterior noise-count COO resident in HBM, sorted by (m, c): m int64, c int32, log_prob fp32, plus the
    non-zero noise-count offsets as parallel arrays sorted by m (the reference's Dict[int, int])."""

    m: torch.Tensor
    c: torch.Tensor
    log_prob: torch.Tensor
    off_m: torch.Tensor
    off_v: torch.Tensor
    n_cols: int  # number of noise-count columns (max c + 1)

    @staticmethod
    def from_scipy(coo: sp.coo_matrix, noise_offsets: Optional[Dict[int, int]], device) -> "DevicePosteriorCOO":
        m = torch.as_tensor(np.asarray(coo.row).astype(np.int64), device=device)
        c = torch.as_tensor(np.asarray(coo.col).astype(np.int32), device=device)
        lp = torch.as_tensor(np.asarray(coo.data).astype(np.float32), device=device)
        if noise_offsets:
            keys = np.fromiter(noise_offsets.keys(), dtype=np.int64, count=len(noise_offsets))
            vals = np.fromiter(noise_offsets.values(), dtype=np.int64, count=len(noise_offsets)).astype(np.int32)
            order = np.argsort(keys, kind="stable")
            off_m = torch.as_tensor(keys[order], device=device)
            off_v = torch.as_tensor(vals[order], device=device)
        else:
            off_m = torch.zeros(0, dtype=torch.int64, device=device)
            off_v = torch.zeros(0, dtype=torch.int32, device=device)
        n_cols = int(c.max().item()) + 1 if c.numel() else 0
        out = DevicePosteriorCOO(m, c, lp, off_m, off_v, n_cols)
        out.ensure_sorted()
        return out

    def ensure_sorted(self):
        """Sort by (m, c) if the COO is not already (the posterior kernel emits it sorted)."""
        n = self.m.numel()
        if n == 0:
            return
        head = torch.empty(n, dtype=torch.int32, device=self.m.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.m.device)
        _cabi.lib().call("cb_segment_heads", ptr(self.m), ptr(self.c), n, ptr(head), ptr(flag), current_stream())
        if int(flag.item()):
            order = torch.argsort(self.c, stable=True)
            order = order[torch.argsort(self.m[order], stable=True)]
            self.m, self.c, self.log_prob = self.m[order].contiguous(), self.c[order].contiguous(), \
                self.log_prob[order].contiguous()

    def to_scipy(self, shape) -> sp.coo_matrix:
        return sp.coo_matrix((self.log_prob.cpu().numpy(), (self.m.cpu().numpy(), self.c.cpu().numpy())), shape=shape)

    def offsets_dict(self) -> Dict[int, int]:
        return dict(zip(self.off_m.cpu().tolist(), self.off_v.cpu().tolist()))


@dataclass
class _Segments:
    n_seg: int
    seg_start: torch.Tensor
    seg_of: torch.Tensor


def _segments(p: DevicePosteriorCOO) -> _Segments:
    n = p.m.numel()
    dev = p.m.device
    L = _cabi.lib()
    head = torch.empty(n, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("cb_segment_heads", ptr(p.m), ptr(p.c), n, ptr(head), ptr(flag), current_stream())
    pos, total = exclusive_scan_i32(head)
    n_seg = int(total.item())
    seg_start = torch.empty(max(n_seg, 1), dtype=torch.int64, device=dev)
    seg_of = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    L.call("cb_segment_starts", ptr(head), ptr(pos), n, ptr(seg_start), ptr(seg_of), current_stream())
    return _Segments(n_seg, seg_start, seg_of)


def _segment_estimate(p: DevicePosteriorCOO, seg: _Segments, mode: int, q: float = 0.5, seed: int = 0,
                      want_map_pos: bool = False):
    dev = p.m.device
    n_seg = seg.n_seg
    seg_m = torch.empty(max(n_seg, 1), dtype=torch.int64, device=dev)
    out_i = torch.empty(max(n_seg, 1), dtype=torch.int32, device=dev)
    out_f = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev) if mode == MODE_MEAN else None
    map_pos = torch.empty(max(n_seg, 1), dtype=torch.int64, device=dev) if want_map_pos else None
    _cabi.lib().call("cb_segment_estimate", mode, ptr(p.m), ptr(p.c), ptr(p.log_prob), ptr(seg.seg_start), n_seg,
                     p.m.numel(), ptr(p.off_m), ptr(p.off_v), p.off_m.numel(), float(q), int(p.n_cols),
                     int(seed) & (2 ** 64 - 1), ptr(seg_m), ptr(out_i), ptr(out_f), ptr(map_pos), current_stream())
    return seg_m[:n_seg], (out_f if mode == MODE_MEAN else out_i)[:n_seg], (map_pos[:n_seg] if want_map_pos else None)


def _segments_to_csr(seg_m: torch.Tensor, values: torch.Tensor, shape, dtype) -> sp.csr_matrix:
    """_estimation_array_to_csr (estimation.py:207-235): rows/cols from m, CSR of the count-matrix shape."""
    n_rows, n_cols = int(shape[0]), int(shape[1])
    n_seg = seg_m.numel()
    if n_seg and (int(seg_m[-1].item()) >= n_rows * n_cols or int(seg_m[0].item()) < 0):
        bad = seg_m[(seg_m < 0) | (seg_m >= n_rows * n_cols)].cpu().numpy()
        raise ValueError(f"Requested m_inds out of range: {bad}")
    dev = seg_m.device
    indptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    col = torch.empty(max(n_seg, 1), dtype=torch.int32, device=dev)
    _cabi.lib().call("cb_csr_from_segments", ptr(seg_m), n_seg, n_rows, n_cols, ptr(indptr), ptr(col), current_stream())
    idx_dtype = np.int32 if n_seg < 2 ** 31 else np.int64
    return sp.csr_matrix((values.cpu().numpy().astype(dtype), col[:n_seg].cpu().numpy(),
                          indptr.cpu().numpy().astype(idx_dtype)), shape=(n_rows, n_cols))


def _as_device_posterior(noise_log_prob_coo, noise_offsets, device) -> DevicePosteriorCOO:
    if isinstance(noise_log_prob_coo, DevicePosteriorCOO):
        return noise_log_prob_coo
    if not torch.cuda.is_available():
        raise RuntimeError("cellbender_b200 estimators need a CUDA device (no CPU fallback)")
    dev = torch.device("cuda" if device in (None, "cpu") else device)
    return DevicePosteriorCOO.from_scipy(noise_log_prob_coo, noise_offsets, dev)


class EstimationMethod(ABC):
    """Base class for estimation of noise counts, given a posterior (estimation.py:29-86)."""

    def __init__(self, index_converter: IndexConverter):
        self.index_converter = index_converter
        super().__init__()

    @abstractmethod
    def estimate_noise(self, noise_log_prob_coo, noise_offsets: Optional[Dict[int, int]], **kwargs) -> sp.csr_matrix:
        ...

    def _simple(self, noise_log_prob_coo, noise_offsets, device, mode, dtype, **kw) -> sp.csr_matrix:
        p = _as_device_posterior(noise_log_prob_coo, noise_offsets, device)
        seg = _segments(p)
        seg_m, val, _ = _segment_estimate(p, seg, mode, **kw)
        return _segments_to_csr(seg_m, val, self.index_converter.matrix_shape, dtype)


class SingleSample(EstimationMethod):
    """A single sample from the noise count posterior (estimation.py:89-114)."""

    @torch.no_grad()
    def estimate_noise(self, noise_log_prob_coo, noise_offsets, device: str = "cuda", seed: Optional[int] = None,
                       **kwargs) -> sp.csr_matrix:
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # consumes the torch CPU generator, as the reference does
        return self._simple(noise_log_prob_coo, noise_offsets, device, MODE_SAMPLE, COUNT_DATATYPE, seed=seed)


class Mean(EstimationMethod):
    """Posterior mean (estimation.py:117-143); float32 output."""

    def estimate_noise(self, noise_log_prob_coo, noise_offsets, device: str = "cuda", **kwargs) -> sp.csr_matrix:
        return self._simple(noise_log_prob_coo, noise_offsets, device, MODE_MEAN, np.float32)


class MAP(EstimationMethod):
    """The canonical maximum a posteriori (estimation.py:146-172)."""

    def estimate_noise(self, noise_log_prob_coo, noise_offsets, device: str = "cuda", **kwargs) -> sp.csr_matrix:
        return self._simple(noise_log_prob_coo, noise_offsets, device, MODE_MAP, COUNT_DATATYPE)


class ThresholdCDF(EstimationMethod):
    """Noise estimation via thresholding the noise count CDF (estimation.py:175-204)."""

    def estimate_noise(self, noise_log_prob_coo, noise_offsets, q: float = 0.5, device: str = "cuda",
                       **kwargs) -> sp.csr_matrix:
        return self._simple(noise_log_prob_coo, noise_offsets, device, MODE_CDF, COUNT_DATATYPE, q=q)


class MultipleChoiceKnapsack(EstimationMethod):
    """Noise estimation via solving a constrained multiple choice knapsack problem (estimation.py:420-702).

    The reference splits genes into ~5000-gene chunks only to bound pandas memory; the chunks are independent
    (estimation.py:531-550) and summed (:529), so one device pass over all genes gives the identical matrix.
    ``n_chunks`` / ``use_multiple_processes`` are accepted and ignored."""

    def estimate_noise(self, noise_log_prob_coo, noise_offsets, noise_targets_per_gene: Optional[np.ndarray] = None,
                       verbose: bool = False, n_chunks: Optional[int] = None, use_multiple_processes: bool = False,
                       device: str = "cuda", **kwargs) -> sp.csr_matrix:
        assert noise_offsets is not None or isinstance(noise_log_prob_coo, DevicePosteriorCOO), \
            "noise_offsets is required for MultipleChoiceKnapsack.estimate_noise"
        assert noise_targets_per_gene is not None, \
            "noise_targets_per_gene is required for MultipleChoiceKnapsack.estimate_noise"
        n_genes = self.index_converter.total_n_genes
        targets = np.asarray(noise_targets_per_gene.detach().cpu().numpy()
                             if isinstance(noise_targets_per_gene, torch.Tensor) else noise_targets_per_gene,
                             dtype=np.float64)
        assert targets.size == n_genes, (f"The number of noise count targets ({targets.size}) "
                                         f"must match the number of genes ({n_genes})")
        p = _as_device_posterior(noise_log_prob_coo, noise_offsets, device)
        seg_m, out = mckp_device(p, targets, n_genes)
        return _segments_to_csr(seg_m, out, self.index_converter.matrix_shape, COUNT_DATATYPE)


@torch.no_grad()
def mckp_device(p: DevicePosteriorCOO, targets: np.ndarray, n_genes: int):
    """MCKP on the device: returns (seg_m int64[n_seg], noise int32[n_seg])."""
    dev = p.m.device
    L = _cabi.lib()
    st = current_stream()
    n = p.m.numel()
    seg = _segments(p)
    n_seg = seg.n_seg
    seg_m, map_val, map_pos = _segment_estimate(p, seg, MODE_MAP, want_map_pos=True)
    tgt = torch.as_tensor(targets, dtype=torch.float64, device=dev)
    gene_sum = torch.empty(n_genes, dtype=torch.int64, device=dev)
    deficit = torch.empty(n_genes, dtype=torch.int64, device=dev)
    L.call("cb_mckp_gene_deficit", ptr(seg_m), ptr(map_val), n_seg, ptr(tgt), n_genes, ptr(gene_sum), ptr(deficit), st)
    steps = torch.zeros(max(n_seg, 1), dtype=torch.int32, device=dev)
    if n:
        is_cand = torch.empty(n, dtype=torch.int32, device=dev)
        key = torch.empty(n, dtype=torch.int64, device=dev)  # uint64 bit pattern; genes < 2^31 keep it non-negative
        L.call("cb_mckp_candidates", ptr(p.m), ptr(p.c), ptr(p.log_prob), ptr(seg.seg_of), ptr(map_pos), ptr(deficit), n,
               n_genes, ptr(is_cand), ptr(key), st)
        pos, total = exclusive_scan_i32(is_cand)
        n_cand = int(total.item())
        if n_cand:
            cand_key = torch.empty(n_cand, dtype=torch.int64, device=dev)
            cand_idx = torch.empty(n_cand, dtype=torch.int64, device=dev)
            L.call("cb_mckp_compact", ptr(is_cand), ptr(pos), ptr(key), n, ptr(cand_key), ptr(cand_idx), st)
            # per gene, smallest delta first; ties keep (m, c) order: STABLE radix sort on (gene << 32 | delta bits)
            sorted_key, perm = torch.sort(cand_key, stable=True)
            sorted_idx = cand_idx[perm].contiguous()
            L.call("cb_mckp_select", ptr(sorted_key), ptr(sorted_idx), n_cand, ptr(deficit), ptr(seg.seg_of), ptr(steps),
                   st)
    out = torch.empty(max(n_seg, 1), dtype=torch.int32, device=dev)
    L.call("cb_mckp_finalize", ptr(seg_m), ptr(map_val), ptr(steps), ptr(deficit), n_seg, n_genes, ptr(out), st)
    return seg_m, out[:n_seg]


def compute_mean_target_removal_as_function(noise_count_posterior_coo, noise_offsets, index_converter: IndexConverter,
                                            raw_count_csr_for_cells: sp.csr_matrix, n_cells: int, device: str,
                                            per_gene: bool):
    """Target removal as a function of FPR (reference posterior.py:1589-1650); same return convention
    (a function of fpr giving the per-cell-scaled target as a torch tensor)."""
    mean_noise_csr = Mean(index_converter=index_converter).estimate_noise(
        noise_log_prob_coo=noise_count_posterior_coo, noise_offsets=noise_offsets, device=device)
    approx_signal_csr = raw_count_csr_for_cells - mean_noise_csr

    def _target_fun(fpr: float) -> torch.Tensor:
        if per_gene:
            target = np.array(mean_noise_csr.sum(axis=0)).squeeze()
            target = target + fpr * np.array(approx_signal_csr.sum(axis=0)).squeeze()
        else:
            target = mean_noise_csr.sum()
            target = target + fpr * approx_signal_csr.sum()
        return torch.tensor(target / n_cells).to(device)

    return _target_fun
