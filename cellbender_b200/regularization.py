"""Posterior regularisation on the GPU behind the reference's interface (reference
cellbender/remove_background/posterior.py:970-1225 ``PosteriorRegularization`` / ``PRq``; ``torch_binary_search``
posterior.py:1654-1746).  SURVEY.md section 8f rank 3: the same segmented-per-m structure as the estimators.

PRq: approximate noise-CDF quantile targeting -- every (cell, gene) posterior p(c) is tilted to p(c) exp(beta n_c) with
the per-entry beta that moves its mean to mean + alpha * std.  The reference densifies ~1 GB chunks and runs a
vectorised bisection until every row of the chunk has converged; here one thread owns one (cell, gene) segment of the
device-resident COO and runs its own bisection (csrc/estimation.cu: prq_regularize_kernel), which gives the same beta
because a converged row's bracket no longer moves.  PRmu / PRmu_gene are not implemented (raise).
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
    """Approximate noise mean targeting (posterior.py:1228-1530).  Not on the B200 path yet."""

    @staticmethod
    def name():
        return "PRmu"

    @staticmethod
    def regularize(noise_count_posterior_coo, noise_offsets, **kwargs):
        raise NotImplementedError("PRmu / PRmu_gene posterior regularisation is not implemented on the B200 path "
                                  "(SURVEY.md section 8f rank 3); PRq is")


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
