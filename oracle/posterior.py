"""ORACLE (test infrastructure, NOT the product path): CPU restatement of the reference's
posterior noise-count log-probability, dense and dtype-generic (float64 = parity target).

Reference locations (relative to /root/reference/cellbender/remove_background):
  Posterior._log_prob_noise_count_tensor   posterior.py:784-859
  Posterior.noise_log_pdf                  posterior.py:669-735
  Posterior._get_cell_noise_count_posterior_coo (thresholding / sparsify)  posterior.py:528-576
  dense_to_sparse_op_torch                 sparse_utils.py:10-33
  IndexConverter                           posterior.py:1533-1586
Third-party arithmetic restated (torch 2.11 torch.distributions, source in-image):
  Poisson.log_prob = xlogy(value, rate) - rate - lgamma(value + 1)
  NegativeBinomial(total_count, logits).log_prob =
      total_count*logsigmoid(-logits) + value*logsigmoid(logits)
      + lgamma(total_count + value) - lgamma(1 + value) - lgamma(total_count)

Pinned by tools/make_golden.py against the unmodified reference static method.
"""
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def log_prob_noise_count_tensor(data, mu_est, lambda_est, alpha_est, n_counts_max: int = 100,
                                mimic_float_casts: bool = True):
    """Un-normalised log prob of noise counts, dense [n, g, c] + window start [n, g].

    posterior.py:815-859.  ``n = min(n_counts_max, data.max())`` is batch-wide.
    Window start ``low = clamp(min(int(lam - n/2), int(x - n + 1)), 0)`` (int() truncates
    toward zero, as torch ``.int()`` does).

    ``mimic_float_casts``: the reference builds the noise-count axis with ``.float()``
    (posterior.py:819,824), so when it is fed float64 inputs lgamma(c + 1) is still
    rounded to float32.  True reproduces that (used to pin against the golden vectors);
    False evaluates the same formula in the input dtype throughout (the parity target
    for the CUDA kernel)."""
    n = int(min(n_counts_max, data.max().item()))
    low = (lambda_est - n / 2).int()
    cdt = torch.float32 if mimic_float_casts else data.dtype
    low = torch.clamp(torch.min(low, (data - n + 1).int()), min=0).to(cdt)
    c = torch.arange(0, n, dtype=cdt, device=data.device).expand(data.shape[0], data.shape[1], n) + low.unsqueeze(-1)
    lam = lambda_est.unsqueeze(-1)
    pois = torch.xlogy(c, lam) - lam - (c + 1).lgamma()
    k = data.unsqueeze(-1) - c
    if alpha_est is None:
        mu = mu_est.unsqueeze(-1)
        other = torch.xlogy(k, mu) - mu - (k + 1).lgamma()
    else:
        alpha_est = torch.as_tensor(alpha_est, dtype=data.dtype, device=data.device)
        logits = (mu_est.log() - alpha_est.log()).unsqueeze(-1)
        tc = alpha_est.unsqueeze(-1) if alpha_est.dim() > 0 else alpha_est
        unnorm = tc * F.logsigmoid(-logits) + k * F.logsigmoid(logits)
        norm = -torch.lgamma(tc + k) + torch.lgamma(1.0 + k) + torch.lgamma(tc)
        norm = norm.masked_fill(tc + k == 0.0, 0.0)
        other = unnorm - norm
    lp = pois + other
    lp = torch.where(c <= data.unsqueeze(-1), lp, torch.full_like(lp, -math.inf))
    return lp, low


def noise_log_pdf(data, samples: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                  n_counts_max: int = 50, lambda_multiplier: float = 1.0):
    """Consensus noise-count log pdf over latent samples (posterior.py:698-735):
    each sample is normalised over c (logsumexp), then a running mean in log space
    ``logaddexp(prev + log(1 - 1/s), cur + log(1/s))``."""
    total, low = None, None
    for s, (mu, lam, alpha) in enumerate(samples, start=1):
        lp, low = log_prob_noise_count_tensor(
            data, mu + 1e-30, lam * lambda_multiplier + 1e-30, alpha + 1e-30, n_counts_max=n_counts_max
        )
        lp = lp - torch.logsumexp(lp, dim=-1, keepdim=True)
        if s == 1:
            total = lp
        else:
            total = torch.logaddexp(
                total + torch.log(torch.tensor(1.0 - 1.0 / s, dtype=data.dtype)),
                lp + torch.log(torch.tensor(1.0 / s, dtype=data.dtype)),
            )
    return total, low


def sparsify_posterior(noise_log_pdf_NGC, low_NG, data, smallest_log_probability: float = -10.0):
    """Dense [n,g,c] -> (n, g, c, log_prob, offset_of_entry) for kept entries, in
    torch.nonzero (row-major) order (posterior.py:531-539, 560-573)."""
    keep = noise_log_pdf_NGC.clone().exp()
    keep[data == 0, :] = 0.0
    keep[noise_log_pdf_NGC < smallest_log_probability] = 0.0
    n_i, g_i, c_i = torch.nonzero(keep, as_tuple=True)
    return n_i, g_i, c_i, noise_log_pdf_NGC[n_i, g_i, c_i], low_NG[n_i, g_i]


def m_index(cell_inds, gene_inds, total_n_genes: int):
    """IndexConverter.get_m_indices (posterior.py:1554-1575): m = n * G_all + g (int64)."""
    return np.asarray(cell_inds).astype(np.int64) * int(total_n_genes) + np.asarray(gene_inds).astype(np.int64)


def posterior_coo_for_batch(data, samples, barcode_inds_all, gene_inds_all, total_n_genes,
                            n_counts_max=20, smallest_log_probability=-10.0):
    """One minibatch of posterior.py:509-576: returns (m, c, log_prob, offsets{m: off!=0}).

    ``barcode_inds_all[b]`` / ``gene_inds_all[g]`` map the batch row / analysed gene to
    indices of the full count matrix."""
    pdf, low = noise_log_pdf(data, samples, n_counts_max=n_counts_max)
    n_i, g_i, c_i, lp, off = sparsify_posterior(pdf, low, data, smallest_log_probability)
    m = m_index(np.asarray(barcode_inds_all)[n_i.numpy()], np.asarray(gene_inds_all)[g_i.numpy()], total_n_genes)
    off = off.numpy().astype(np.int64)
    offsets: Dict[int, int] = {int(k): int(v) for k, v in zip(m[off != 0], off[off != 0])}
    return m, c_i.numpy().astype(np.int32), lp.numpy(), offsets


# ----------------------------------------------------------------------------------------------------
# PRq posterior regularisation (posterior.py:1002-1225) and torch_binary_search (posterior.py:1654-1746),
# restated per m-segment in numpy.  Pinned by tests/golden/prq.npz (unmodified reference, tools/make_golden.py).
# dtype flow of the reference: dense log prob float64; counts n_c = c + offset, log n_c and beta * n_c float32;
# bisection bracket float32; target log(mean + alpha std) float64 with float32 column indices.
# ----------------------------------------------------------------------------------------------------
def prq_regularize(rows, cols, data, offsets: Dict[int, int], alpha: float, target_tolerance: float = 0.001,
                   max_iterations: int = 100):
    """Returns (m, c, log_prob_reg float64) sorted by (m, c), and {m: log target}."""
    rows, cols = np.asarray(rows).astype(np.int64), np.asarray(cols).astype(np.int64)
    data = np.asarray(data).astype(np.float64)
    order = np.lexsort((cols, rows))
    rows, cols, data = rows[order], cols[order], data[order]
    out_m, out_c, out_lp, targets = [], [], [], {}
    starts = np.flatnonzero(np.r_[True, rows[1:] != rows[:-1]])
    ends = np.r_[starts[1:], rows.size]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for a, b in zip(starts, ends):
            m = int(rows[a])
            c, lp = cols[a:b], data[a:b]
            prob = np.exp(lp)
            cf = c.astype(np.float32).astype(np.float64)
            mean = float((cf * prob).sum())
            std = float(np.sqrt((((cf - mean) ** 2) * prob).sum()))
            target = np.log(np.float64(mean + alpha * std))
            targets[m] = float(target)
            n32 = (c + offsets.get(m, 0)).astype(np.float32)
            logn = np.log(n32).astype(np.float64)  # float32 log, promoted
            lo, hi, beta = np.float32(-100.0), np.float32(100.0), np.float32(0.0)
            for _ in range(max_iterations):
                beta = np.float32((lo + hi) * np.float32(0.5))
                td = lp + (beta * n32).astype(np.float64)
                outcome = _lse(logn + td) - _lse(td) - target
                if abs(0.0 - outcome) < target_tolerance:
                    break
                if outcome < -target_tolerance:
                    lo = beta
                if outcome > target_tolerance:
                    hi = beta
            reg = lp + (beta * n32).astype(np.float64)
            reg = reg - _lse(reg)
            keep = np.exp(reg) != 0
            out_m += [m] * int(keep.sum())
            out_c += c[keep].tolist()
            out_lp += reg[keep].tolist()
    return np.array(out_m, dtype=np.int64), np.array(out_c, dtype=np.int32), np.array(out_lp, dtype=np.float64), targets


def _lse(v):
    mx = np.max(v)
    if not np.isfinite(mx):
        return mx
    return mx + np.log(np.exp(v - mx).sum())
