"""ORACLE (test infrastructure, NOT the product path): CPU restatement of the reference's
rate mixtures and NB (+) Poisson likelihoods in plain torch, dtype-generic (run it in
float64 for the parity target, float32 to reproduce the reference's own rounding).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

Pinned against the reference itself: ``tools/make_golden.py`` imports the unmodified
reference functions (through ``oracle/ref_shim.py``) and stores their outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.

Reference locations (relative to /root/reference):
  calculate_lambda   cellbender/remove_background/model.py:25-58
  calculate_mu       cellbender/remove_background/model.py:61-85
  NBPCapprox.log_prob  cellbender/remove_background/distributions/NegativeBinomialPoissonConvApprox.py:105-143
  NBPC.log_prob        cellbender/remove_background/distributions/NegativeBinomialPoissonConv.py:89-154
"""
from typing import Optional

import torch


def calculate_lambda(model_type: str, epsilon, chi_ambient, d_empty, y=None, d_cell=None, rho=None, chi_bar=None):
    """Noise rate lambda[.., g] (model.py:25-58)."""
    e = epsilon.unsqueeze(-1)
    de = d_empty.unsqueeze(-1)
    if model_type in ("simple", "ambient"):
        return e * de * chi_ambient
    if model_type == "swapping":
        return rho.unsqueeze(-1) * chi_bar * e * (y.unsqueeze(-1) * d_cell.unsqueeze(-1) + de)
    if model_type == "full":
        r = rho.unsqueeze(-1)
        return e * ((1.0 - r) * chi_ambient * de + r * chi_bar * (y.unsqueeze(-1) * d_cell.unsqueeze(-1) + de))
    raise NotImplementedError(model_type)


def calculate_mu(model_type: str, epsilon, d_cell, chi, y=None, rho=None):
    """Cell-count mean mu[.., g] (model.py:61-85)."""
    e = epsilon.unsqueeze(-1)
    dc = d_cell.unsqueeze(-1)
    if model_type == "simple":
        return e * dc * chi
    if model_type == "ambient":
        return y.unsqueeze(-1) * e * dc * chi
    if model_type in ("swapping", "full"):
        return (1.0 - rho.unsqueeze(-1)) * y.unsqueeze(-1) * e * dc * chi
    raise NotImplementedError(model_type)


def _poisson_log_prob(lam, value):
    return lam.log() * value - lam - (value + 1).lgamma()


def _neg_binom_log_prob(mu, alpha, value):
    return (
        (value + alpha).lgamma()
        - (value + 1).lgamma()
        - alpha.lgamma()
        + alpha * (alpha.log() - (alpha + mu).log())
        + value * (mu.log() - (alpha + mu).log())
    )


def nbpc_approx_log_prob(value, mu, alpha, lam):
    """Moment-matched NB approximation of NB(mu, alpha) (+) Poisson(lam); Poisson(lam)
    where mu < 1e-5 * lam  (...ConvApprox.py:105-143, formula lines 63-70, 117-127).
    The safeguard epsilons (model.py:338-346) are the CALLER's job, as in the reference."""
    mu, alpha, lam, value = torch.broadcast_tensors(mu, torch.as_tensor(alpha, dtype=mu.dtype), lam, value)
    mean_hat = mu + lam
    alpha_hat = mean_hat.pow(2) * alpha * mu.pow(-2)
    nb = _neg_binom_log_prob(mean_hat, alpha_hat, value)
    empty = mu < 1e-5 * lam
    return torch.where(empty, _poisson_log_prob(lam, value), nb)


def nbpc_exact_log_prob(value, mu, alpha, lam, max_poisson: int = 100):
    """Exact NB (+) Poisson convolution by direct summation (...Conv.py:89-154):
    closed forms for value 0 and 1, otherwise logsumexp over max_poisson+1 Poisson terms
    starting at clamp(min(int(lam - max_poisson/2), int(value - max_poisson)), 0)."""
    mu, alpha, lam, value = torch.broadcast_tensors(mu, torch.as_tensor(alpha, dtype=mu.dtype), lam, value)
    log_a_ratio = alpha.log() - (alpha + mu).log()
    lp_zero = -lam + alpha * log_a_ratio
    nb_one = (alpha + 1) * log_a_ratio + mu.log()
    lp_one = torch.logaddexp(-lam + nb_one, (lam.log() - lam) + alpha * log_a_ratio)

    low = (lam.detach() - max_poisson / 2).int()
    low = torch.clamp(torch.min(low, (value - max_poisson).int()), min=0)
    inc = torch.arange(0, max_poisson + 1, dtype=value.dtype, device=value.device)
    pois_vals = low.unsqueeze(-1).to(value.dtype) + inc
    nb_vals = value.unsqueeze(-1) - pois_vals
    bad = nb_vals < 0
    nb_vals = nb_vals.clamp(min=0)
    # NOTE the reference casts the count tables with ``.float()`` (...Conv.py:138-141), so
    # their lgamma terms are evaluated in float32 even when mu/alpha/lam are float64.
    tot = _poisson_log_prob(lam.unsqueeze(-1), pois_vals.float()) + _neg_binom_log_prob(
        mu.unsqueeze(-1), alpha.unsqueeze(-1), nb_vals.float()
    )
    tot = torch.where(bad, torch.full_like(tot, float("-inf")), tot)
    lp_rest = torch.logsumexp(tot, dim=-1)
    return torch.where(value == 0, lp_zero, torch.where(value == 1, lp_one, lp_rest))


def get_p_logit_prior(log_counts, log_cell_prior_counts, surely_empty_counts, naive_p_logit_prior):
    """Per-droplet logit prior of containing a cell (model.py:577-592)."""
    ones = torch.ones_like(log_counts)
    out = ones * naive_p_logit_prior
    thresh = (torch.as_tensor(surely_empty_counts, dtype=log_counts.dtype).log() + log_cell_prior_counts) / 2
    out = torch.where(log_counts <= thresh, ones * -100.0, out)
    out = torch.where(log_counts >= log_cell_prior_counts, ones * 10.0, out)
    return out
