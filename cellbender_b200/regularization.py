"""Posterior regularisation on the GPU behind the reference's interface (reference
cellbender/remove_background/posterior.py:970-1225 ``PosteriorRegularization`` / ``PRq``; ``torch_binary_search``
posterior.py:1654-1746).  SURVEY.md section 8f rank 3: the same segmented-per-m structure as the estimators.

PRq: approximate noise-CDF quantile targeting -- every (cell, gene) posterior p(c) is tilted to p(c) exp(beta n_c) with
the per-entry beta that moves its mean to mean + alpha * std.  The reference densifies ~1 GB chunks and runs a
vectorised bisection until every row of the chunk has converged; here one thread owns one (cell, gene) segment of the
device-resident COO and runs its own bisection (csrc/estimation.cu: prq_regularize_kernel), which gives the same beta
because a converged row's bracket no longer moves.
PRmu / PRmu_gene: approximate mean targeting -- one overall (or per-gene) tilt factor found by bisection on the MAP noise
counts of a random subset of cells; the tilt, the MAP and the per-gene sums all run on the device-resident COO.
"""
from abc import ABC, abstractmethod
from typing import Callable, Dict, Optional

import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi, consts
from ._cabi import current_stream, ptr
from .estimation import DevicePosteriorCOO, _as_device_posterior, _segments

POSTERIOR_REG_SEARCH_MAX_ITER = 100  # consts.py:96


class PosteriorRegularization(ABC):
    """Base class (posterior.py:980-1000)."""

    def __init__(self):
        super().__init__()

    @staticmethod
    @abstractmethod
    def name():
        """Short name of this regularization method"""

    @staticmethod
    @abstractmethod
    def regularize(noise_count_posterior_coo, noise_offsets: Optional[Dict[int, int]], **kwargs):
        """Perform posterior regularization"""


def _prq_device(p: DevicePosteriorCOO, alpha: float, target_tolerance: float, max_iterations: int):
    """Run the PRq kernel over a device posterior.  Returns (seg, seg_m, log_target fp64 [n_seg], beta fp32 [n_seg],
    regularised log prob fp64 [n_entries], keep int32 [n_entries])."""
    seg = _segments(p)
    n, n_seg, dev = p.m.numel(), seg.n_seg, p.m.device
    log_target = torch.empty(max(n_seg, 1), dtype=torch.float64, device=dev)
    beta = torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev)
    lp_out = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    keep = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    _cabi.lib().call("cb_prq_regularize", ptr(p.m), ptr(p.c), ptr(p.log_prob), ptr(seg.seg_start), n_seg, n, ptr(p.off_m),
                     ptr(p.off_v), p.off_m.numel(), float(alpha), float(target_tolerance), int(max_iterations),
                     ptr(log_target), ptr(beta), ptr(lp_out), ptr(keep), current_stream())
    seg_m = p.m[seg.seg_start[:n_seg]] if n_seg else p.m[:0]
    return seg, seg_m, log_target[:n_seg], beta[:n_seg], lp_out[:n], keep[:n]


class PRq(PosteriorRegularization):
    """Approximate noise CDF quantile targeting: E_reg[noise] >= E[noise] + alpha * Std[noise] (posterior.py:1002-1225)."""

    @staticmethod
    def name():
        return "PRq"

    @staticmethod
    def _compute_log_target_dict(noise_count_posterior_coo, alpha: float, device: str = "cuda") -> Dict[int, float]:
        """log(mean + alpha * std) for each 'm' index (posterior.py:1021-1042)."""
        p = _as_device_posterior(noise_count_posterior_coo, None, device)
        _, seg_m, log_target, _, _, _ = _prq_device(p, alpha, 1e-3, 1)
        return dict(zip(seg_m.cpu().tolist(), log_target.cpu().tolist()))

    @staticmethod
    @torch.no_grad()
    def regularize_device(posterior: DevicePosteriorCOO, alpha: float = 0.0, target_tolerance: float = 0.001,
                          max_iterations: int = POSTERIOR_REG_SEARCH_MAX_ITER) -> DevicePosteriorCOO:
        """The regularised posterior, computed and left on the device (log prob rounded to fp32, the estimators' input
        type; offsets carried over)."""
        assert target_tolerance > 0, "target_tolerance should be > 0."
        _, _, _, _, lp, keep = _prq_device(posterior, alpha, target_tolerance, max_iterations)
        mask = keep.bool()
        return DevicePosteriorCOO(posterior.m[mask].contiguous(), posterior.c[mask].contiguous(),
                                  lp[mask].float().contiguous(), posterior.off_m, posterior.off_v, posterior.n_cols)

    @staticmethod
    @torch.no_grad()
    def regularize(noise_count_posterior_coo: sp.coo_matrix, noise_offsets: Optional[Dict[int, int]], alpha: float = 0.0,
                   device: str = "cuda", target_tolerance: float = 0.001, n_chunks: Optional[int] = None,
                   **kwargs) -> sp.coo_matrix:
        """Same signature and return type as the reference (posterior.py:1177-1225): a scipy COO of the regularised log
        probabilities (float64, as the reference's python-float lists give) with the input's shape, entries in (m, c)
        order.  ``n_chunks`` is accepted and ignored (segments are independent)."""
        assert target_tolerance > 0, "target_tolerance should be > 0."
        p = _as_device_posterior(noise_count_posterior_coo, noise_offsets, device)
        _, _, _, _, lp, keep = _prq_device(p, alpha, target_tolerance, POSTERIOR_REG_SEARCH_MAX_ITER)
        mask = keep.bool()
        return sp.coo_matrix((lp[mask].cpu().numpy(), (p.m[mask].cpu().numpy(), p.c[mask].cpu().numpy())),
                             shape=noise_count_posterior_coo.shape)


class PRmu(PosteriorRegularization):
    r"""Approximate noise mean targeting (posterior.py:1228-1530): one tilt factor beta (overall, the v0.2.0
    behaviour) or one per gene, found by bisection so that the MAP noise counts of the regularised posterior, summed over
    a random subset of cells, equal E[noise] + nFPR * raw counts."""

    @staticmethod
    def name():
        return "PRmu"

    @staticmethod
    def _tilt_device(p: DevicePosteriorCOO, beta: torch.Tensor, total_n_genes: int):
        """(regularised log prob fp64 [n_entries], keep mask) of ``p`` tilted by ``beta`` ([1] or [total_n_genes])."""
        seg = _segments(p)
        n, dev = p.m.numel(), p.m.device
        beta = beta.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        lp_out = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
        keep = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _cabi.lib().call("cb_posterior_tilt", ptr(p.m), ptr(p.c), ptr(p.log_prob), ptr(seg.seg_start), seg.n_seg, n, ptr(p.off_m),
                         ptr(p.off_v), p.off_m.numel(), ptr(beta), beta.numel(), int(total_n_genes), ptr(lp_out), ptr(keep),
                         current_stream())
        return lp_out[:n], keep[:n].bool()

    @staticmethod
    def _chunked_compute_regularized_posterior(noise_count_posterior_coo, noise_offsets, index_converter, beta: torch.Tensor,
                                               device: str = "cuda", n_chunks: Optional[int] = None) -> sp.coo_matrix:
        """posterior.py:1305-1393 (chunks are independent: one pass)."""
        p = _as_device_posterior(noise_count_posterior_coo, noise_offsets, device)
        lp, mask = PRmu._tilt_device(p, beta, index_converter.total_n_genes)
        return sp.coo_matrix((lp[mask].cpu().numpy(), (p.m[mask].cpu().numpy(), p.c[mask].cpu().numpy())),
                             shape=noise_count_posterior_coo.shape)

    @staticmethod
    def _subset_posterior_by_cells(p: DevicePosteriorCOO, index_converter, n_cells: int) -> DevicePosteriorCOO:
        """A random slice of the posterior with ``n_cells`` cells (posterior.py:1395-1427; numpy's global RNG chooses
        the cells exactly as in the reference)."""
        G = index_converter.total_n_genes
        cell_of = torch.div(p.m, G, rounding_mode="floor")
        unique_cell_inds = torch.unique(cell_of).cpu().numpy()
        n_cells = min(n_cells, len(unique_cell_inds))
        chosen = np.random.choice(unique_cell_inds, size=n_cells, replace=False)
        lut = torch.zeros(index_converter.total_n_cells, dtype=torch.bool, device=p.m.device)
        lut[torch.as_tensor(chosen, device=p.m.device)] = True
        mask = lut[cell_of]
        return DevicePosteriorCOO(p.m[mask].contiguous(), p.c[mask].contiguous(), p.log_prob[mask].contiguous(), p.off_m,
                                  p.off_v, p.n_cols)

    @staticmethod
    def _binary_search_for_posterior_regularization_factor(p: DevicePosteriorCOO, index_converter, target_removal: torch.Tensor,
                                                           shape: int, target_tolerance: float = 100, max_iterations: int = 20,
                                                           device: str = "cuda") -> torch.Tensor:
        """posterior.py:1245-1303: MAP noise counts of the tilted subset posterior, summed overall or per gene, as a
        function of beta; bisection on [-100, 200]."""
        from .estimation import MODE_MAP, _segment_estimate

        G = index_converter.total_n_genes
        per_gene = target_removal.dim() > 0 and target_removal.numel() > 1

        def summarize_map_noise_counts(x: torch.Tensor) -> torch.Tensor:
            lp, mask = PRmu._tilt_device(p, x, G)
            reg = DevicePosteriorCOO(p.m[mask].contiguous(), p.c[mask].contiguous(), lp[mask].float().contiguous(), p.off_m,
                                     p.off_v, p.n_cols)
            seg_m, map_val, _ = _segment_estimate(reg, _segments(reg), MODE_MAP)
            if per_gene:
                return torch.zeros(G, dtype=torch.int64, device=p.m.device).index_add_(0, seg_m % G, map_val.long())
            return map_val.long().sum()

        dev = p.m.device
        return torch_binary_search(
            evaluate_outcome_given_value=summarize_map_noise_counts, target_outcome=target_removal.to(dev),
            init_range=torch.tensor([-100.0, 200.0], device=dev).unsqueeze(0).expand((shape, 2)),
            target_tolerance=target_tolerance, max_iterations=max_iterations, debug=True)

    @staticmethod
    @torch.no_grad()
    def regularize(noise_count_posterior_coo: sp.coo_matrix, noise_offsets: Optional[Dict[int, int]], index_converter=None,
                   raw_count_matrix: Optional[sp.csr_matrix] = None, fpr: float = 0.01, per_gene: bool = False,
                   device: str = "cuda", target_tolerance: float = 0.5, n_cells: int = 1000, n_chunks: Optional[int] = None,
                   **kwargs) -> sp.coo_matrix:
        """Same signature and return type as the reference (posterior.py:1429-1530)."""
        from .estimation import compute_mean_target_removal_as_function
        from .sparse_utils import combine_csr

        assert index_converter is not None, "index_converter is required for PRmu.regularize"
        assert raw_count_matrix is not None, "raw_count_matrix is required for PRmu.regularize"
        p = _as_device_posterior(noise_count_posterior_coo, noise_offsets, device)
        subset = PRmu._subset_posterior_by_cells(p, index_converter, n_cells)
        G = index_converter.total_n_genes
        included = np.unique(torch.div(subset.m, G, rounding_mode="floor").cpu().numpy())
        excluded = np.setdiff1d(np.arange(raw_count_matrix.shape[0]), included)
        raw_count_csr_for_cells = combine_csr(raw_count_matrix, zero_rows=excluded)  # does not touch the caller's matrix
        target_fun = compute_mean_target_removal_as_function(
            noise_count_posterior_coo=subset, noise_offsets=None, index_converter=index_converter,
            raw_count_csr_for_cells=raw_count_csr_for_cells, n_cells=len(included), device=str(p.m.device), per_gene=per_gene)
        target_removal = target_fun(fpr) * len(included)
        shape = G if per_gene else 1
        beta = PRmu._binary_search_for_posterior_regularization_factor(
            subset, index_converter, target_removal=target_removal, device=device, target_tolerance=target_tolerance,
            shape=shape)
        lp, mask = PRmu._tilt_device(p, beta, G)
        return sp.coo_matrix((lp[mask].cpu().numpy(), (p.m[mask].cpu().numpy(), p.c[mask].cpu().numpy())),
                             shape=noise_count_posterior_coo.shape)


def torch_binary_search(evaluate_outcome_given_value: Callable[[torch.Tensor], torch.Tensor], target_outcome: torch.Tensor,
                        init_range: torch.Tensor, target_tolerance: float = 0.001,
                        max_iterations: int = POSTERIOR_REG_SEARCH_MAX_ITER, debug: bool = False) -> torch.Tensor:
    """Vectorised bisection with the reference's exact update rule (posterior.py:1654-1746): evaluate at the bracket
    midpoint; stop when every residual is within tolerance; otherwise the lower end moves where the outcome is below
    target - tol and the upper end where it is above target + tol.  ``evaluate_outcome_given_value`` must increase
    monotonically.  Works on any device; the PRq kernel inlines this rule per segment."""
    assert target_tolerance > 0, "target_tolerance should be > 0."
    assert len(init_range.shape) > 1, \
        "init_range must be at least two-dimensional (last dimension contains lower and upper bounds)"
    assert init_range.shape[-1] == 2, "Last dimension of init_range should be 2: low and high"
    bracket = init_range.clone()
    value = bracket.mean(dim=-1)
    for _ in range(max_iterations):
        value = bracket.mean(dim=-1)
        outcome = evaluate_outcome_given_value(value)
        if bool(((target_outcome - outcome).abs() < target_tolerance).all()):
            break
        bracket[..., 0] = torch.where(outcome < target_outcome - target_tolerance, value, bracket[..., 0])
        bracket[..., 1] = torch.where(outcome > target_outcome + target_tolerance, value, bracket[..., 1])
    return value
