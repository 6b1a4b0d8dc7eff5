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


def nbpc_exact_log_prob(value, mu, alpha, lam, max_poisson: int = 100, float32_tables: bool = True):
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
    # float32_tables=False: the same sum without that cast (the mathematically exact value in the operands' dtype)
    cast = (lambda t: t.float()) if float32_tables else (lambda t: t)
    tot = _poisson_log_prob(lam.unsqueeze(-1), cast(pois_vals)) + _neg_binom_log_prob(
        mu.unsqueeze(-1), alpha.unsqueeze(-1), cast(nb_vals)
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


# ----------------------------------------------------------------------------------------------------
# Multiprecision truth (small cases only).  The reference formula cancels catastrophically even in
# float64 when alpha_hat = (m/mu)^2 alpha is large (errors up to 1e-3 absolute, gradients pure noise:
# see tests/golden/nbpc.npz 'approx_f64' vs 'approx_mp'), so the CUDA kernels are judged against an
# 60-digit evaluation of the SAME formula and of its analytic derivatives.
# ----------------------------------------------------------------------------------------------------
def nbpc_approx_mp(value, mu, alpha, lam, dps: int = 60):
    """Elementwise (numpy arrays, broadcast beforehand) -> (L, dL/dmu, dL/dlam, dL/dalpha) as float64.
    The Poisson-branch test ``mu < 1e-5 * lam`` is evaluated in float32, as the product evaluates it."""
    import mpmath as mp
    import numpy as np

    value, mu, alpha, lam = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (value, mu, alpha, lam)))
    shape = value.shape
    value, mu, alpha, lam = (a.ravel() for a in (value, mu, alpha, lam))
    L, dmu, dlam, dal = (np.zeros(value.size) for _ in range(4))
    with mp.workdps(dps):
        for i in range(value.size):
            v, m_, a_, l_ = mp.mpf(value[i]), mp.mpf(mu[i]), mp.mpf(alpha[i]), mp.mpf(lam[i])
            if np.float32(mu[i]) < np.float32(1e-5) * np.float32(lam[i]):
                L[i] = float(v * mp.log(l_) - l_ - mp.loggamma(v + 1))
                dlam[i] = float(v / l_ - 1)
                continue
            mh = m_ + l_
            ah = mh ** 2 * a_ / m_ ** 2
            L[i] = float(mp.loggamma(v + ah) - mp.loggamma(v + 1) - mp.loggamma(ah)
                         + ah * (mp.log(ah) - mp.log(ah + mh)) + v * (mp.log(mh) - mp.log(ah + mh)))
            dLda = mp.digamma(v + ah) - mp.digamma(ah) + mp.log(ah / (ah + mh)) + (mh - v) / (ah + mh)
            dLdm = v / mh - (ah + v) / (ah + mh)
            dmu[i] = float(dLdm + dLda * (-2 * ah * l_ / (mh * m_)))
            dlam[i] = float(dLdm + dLda * (2 * ah / mh))
            dal[i] = float(dLda * ah / a_)
    return L.reshape(shape), dmu.reshape(shape), dlam.reshape(shape), dal.reshape(shape)


def elbo_obs_oracle(x, logits, chi_ambient, chi_bar, coef, alpha, w):
    """Truth for the fused observation-term kernel (cb_elbo_obs_fwd_bwd): numpy float64 + multiprecision
    likelihood.  x [B,G] dense counts; coef [B,7] = a_mu1, cA1, cB1, cA0, cB0, cAE, cBE; w [B,3].
    Rates as in model.py:49-54,78-80,338-346,385-397 written in factored form.  Returns a dict with the same
    fields as the kernel: sums [B,12], dlogits [B,G], dchi_ambient [G], lse [B]."""
    import numpy as np

    x, logits, chi_ambient, chi_bar, coef, w = (np.asarray(a, dtype=np.float64) for a in
                                                (x, logits, chi_ambient, chi_bar, coef, w))
    alpha = float(alpha)
    mx = logits.max(-1, keepdims=True)
    lse = (mx + np.log(np.exp(logits - mx).sum(-1, keepdims=True)))
    chi = np.exp(logits - lse)
    a_mu1, cA1, cB1, cA0, cB0, cAE, cBE = (coef[:, i:i + 1] for i in range(7))
    mu1 = a_mu1 * chi + 1e-10
    lam1 = cA1 * chi_ambient + cB1 * chi_bar + 1e-10
    lam0 = cA0 * chi_ambient + cB0 * chi_bar + 1e-10
    lamE = cAE * chi_ambient + cBE * chi_bar + 1e-10
    # the product computes rates in fp32: evaluate the truth at the fp32-rounded rates' float64 values? No --
    # the truth uses exact float64 rates; rate rounding is part of the kernel's error budget.
    L1, dmu1, dlam1, dal1 = nbpc_approx_mp(x, mu1, alpha, lam1)
    L0, _, dlam0, dal0 = nbpc_approx_mp(x, np.full_like(mu1, 1e-10), alpha, lam0)
    from scipy.special import gammaln

    LE = x * np.log(lamE) - lamE - gammaln(x + 1)
    dE = x / lamE - 1.0
    sums = np.stack([
        L1.sum(-1), L0.sum(-1), LE.sum(-1), (dmu1 * chi).sum(-1), (dlam1 * chi_ambient).sum(-1),
        (dlam1 * chi_bar).sum(-1), dal1.sum(-1), (dlam0 * chi_ambient).sum(-1), (dlam0 * chi_bar).sum(-1),
        dal0.sum(-1), (dE * chi_ambient).sum(-1), (dE * chi_bar).sum(-1)], axis=1)
    w1, w0, wE = (w[:, i:i + 1] for i in range(3))
    gchi = w1 * a_mu1 * dmu1
    dlogits = chi * (gchi - (gchi * chi).sum(-1, keepdims=True))
    dchi_amb = (w1 * cA1 * dlam1 + w0 * cA0 * dlam0 + wE * cAE * dE).sum(0)
    return {"sums": sums, "dlogits": dlogits, "dchi_ambient": dchi_amb, "lse": lse[:, 0]}
