"""ORACLE (test infrastructure, NOT the product path): CPU restatement, in plain dense torch, of the ELBO that
pyro's ``TraceEnum_ELBO(max_plate_nesting=1)`` computes for the reference's ``model``/``guide`` pair with ``y``
enumerated in the guide.

PARITY UNPINNED against pyro itself: pyro-ppl (>=1.8.4, unpinned in the reference's requirements.txt:5) is not
installed in this image and no reference test pins an ELBO value (SURVEY.md finding 1, section 4).  What IS pinned:
every building block used here -- rate mixtures, likelihood, encoder/decoder forwards, p_logit prior -- is checked
against the unmodified reference in tests/test_oracle_golden.py; the enumeration / Dice / mask / scale algebra is the
hand restatement of SURVEY.md section 3.2 (model.py:232-445 model, :447-574 guide).

Two likelihood back-ends:
  'reference'  the reference's own formula (oracle/nbpc.py:nbpc_approx_log_prob) in the tensors' dtype -- this is
               what the reference computes, rounding errors included; used for the timed CPU baseline;
  'mp'         multiprecision value + analytic gradient per element (small cases) -- the parity target;
  'exact'      the exact NB (+) Poisson convolution (NegativeBinomialPoissonConv, max_poisson = 100) in the tensors' dtype
               without the reference's float32 cast of its count tables (true value in float64).
"""
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Beta, Gamma, LogNormal, Normal, constraints, transform_to

from . import nbpc as onb
from . import vae as ovae

P_LOGIT_SCALE = 2.0  # consts.py:65
REG_SCALE_AMBIENT_EXPRESSION = 0.01  # consts.py:80
REG_SCALE_SOFT_SUPERVISION = 10.0  # consts.py:81
REG_LOGIT_MEAN = 10.0  # consts.py:84
REG_LOGIT_SOFT_SCALE = 1.0  # consts.py:86
EPS = 1e-10  # consts.py:74-77

CONSTRAINTS = {
    "chi_ambient": constraints.simplex, "d_cell_scale": constraints.positive, "d_empty_loc": constraints.positive,
    "d_empty_scale": constraints.positive, "rho_alpha": constraints.interval(0.001, 1000.0),
    "rho_beta": constraints.interval(0.001, 1000.0), "phi_loc": constraints.positive, "phi_scale": constraints.positive,
}


# ---- reparameterised sampling from FIXED base noise, so that product and oracle see the same randomness ---------
class _GammaFromStd(torch.autograd.Function):
    """standard-gamma sample `g` (drawn at the current concentration) with torch's implicit reparameterisation
    gradient d g / d conc = torch._standard_gamma_grad(conc, g) (what Gamma.rsample differentiates through)."""

    @staticmethod
    def forward(ctx, conc, g):
        ctx.save_for_backward(conc, g)
        return g.clone()

    @staticmethod
    def backward(ctx, grad):
        conc, g = ctx.saved_tensors
        return grad * torch._standard_gamma_grad(conc, g), None


class _BetaFromSample(torch.autograd.Function):
    """Beta(a, b) sample with the Dirichlet implicit gradient torch's Beta.rsample uses (torch/distributions/dirichlet.py)."""

    @staticmethod
    def forward(ctx, a, b, x):
        ctx.save_for_backward(a, b, x)
        return x.clone()

    @staticmethod
    def backward(ctx, grad):
        a, b, x = ctx.saved_tensors
        conc = torch.stack([a.expand_as(x), b.expand_as(x)], -1)
        xx = torch.stack([x, 1 - x], -1)
        go = torch.stack([grad, torch.zeros_like(grad)], -1)
        total = conc.sum(-1, True).expand_as(conc)
        g = torch._dirichlet_grad(xx, conc, total)
        gc = g * (go - (xx * go).sum(-1, True))
        return gc[..., 0].sum_to_size(a.shape), gc[..., 1].sum_to_size(b.shape), None


def gamma_rsample(conc, rate, std_gamma):
    return _GammaFromStd.apply(conc, std_gamma) / rate


def beta_rsample(a, b, x):
    return _BetaFromSample.apply(a, b, x)


def draw_noise(B: int, Z: int, conc: Dict[str, torch.Tensor], generator=None, device="cpu", dtype=torch.float32):
    """Base noise for one minibatch.  ``conc`` = current concentrations {'phi': [], 'epsilon': [B], 'rho_alpha': [],
    'rho_beta': []} at which the gamma / beta samples are drawn."""
    rn = lambda *s: torch.randn(*s, generator=generator, device=device, dtype=dtype)
    sg = lambda c: torch._standard_gamma(c.to(device=device, dtype=dtype)).clamp_min(torch.finfo(dtype).tiny)
    a, b = sg(conc["rho_alpha"].expand(B)), sg(conc["rho_beta"].expand(B))
    return {"z": rn(B, Z), "d_cell": rn(B), "d_empty": rn(B), "p_logit_reg": rn(B), "phi": sg(conc["phi"]),
            "epsilon": sg(conc["epsilon"]), "rho": a / (a + b)}


class _NbpcMp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, mu, alpha, lam):
        v, m, a, l = torch.broadcast_tensors(value, mu, alpha, lam)
        L, dmu, dlam, dal = onb.nbpc_approx_mp(v.detach().numpy(), m.detach().numpy(), a.detach().numpy(), l.detach().numpy())
        ctx.grads = tuple(torch.tensor(g, dtype=mu.dtype) for g in (dmu, dal, dlam))
        ctx.shapes = (mu.shape, alpha.shape, lam.shape)
        return torch.tensor(L, dtype=mu.dtype)

    @staticmethod
    def backward(ctx, g):
        dmu, dal, dlam = ctx.grads
        return None, (g * dmu).sum_to_size(ctx.shapes[0]), (g * dal).sum_to_size(ctx.shapes[1]), (g * dlam).sum_to_size(ctx.shapes[2])


def _masked_median(values, mask):
    return values[mask].median() if bool(mask.any()) else torch.ones((), dtype=values.dtype, device=values.device)


def elbo(x: torch.Tensor, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor], cfg: Dict, noise: Dict,
         row_keep_scale: Optional[torch.Tensor], training: bool, likelihood: str = "reference",
         model_type: str = "full", update_bn: bool = False) -> Dict[str, torch.Tensor]:
    """All ELBO terms for the dense minibatch ``x`` [B, G].

    params: 'enc_z.*', 'enc_o.*', 'dec.*' (reference state_dict keys) and 'ps.<name>' = UNCONSTRAINED pyro params.
    buffers: BatchNorm running statistics under the same prefixes.  cfg: priors and encoder constants.
    noise: see draw_noise.  update_bn: training mode also updates the BatchNorm running statistics in ``buffers``
    (the buffers of ``sub(...)`` views are the same tensors).  Returns the terms (same names as cellbender_b200.model.elbo_terms) and 'loss' = -ELBO."""
    dt = x.dtype
    sd = {**params, **buffers}
    ps = {k[3:]: transform_to(CONSTRAINTS[k[3:]])(v) for k, v in params.items() if k.startswith("ps.")}
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    dev = x.device
    c = lambda v: torch.as_tensor(v, dtype=dt, device=dev)
    B = x.shape[0]
    chi_ambient = ps["chi_ambient"]
    chi_bar = c(cfg["chi_bar"])

    # ---------------- guide (model.py:447-574) ----------------
    phi_conc, phi_rate = ps["phi_loc"].pow(2) / ps["phi_scale"].pow(2), ps["phi_loc"] / ps["phi_scale"].pow(2)
    phi = gamma_rsample(phi_conc, phi_rate, noise["phi"])
    include_rho = model_type in ("full", "swapping")  # model.py:140-141: rho exists only for these model types
    rho = beta_rsample(ps["rho_alpha"], ps["rho_beta"], noise["rho"]) if include_rho else torch.zeros(x.shape[0], dtype=dt, device=x.device)
    d_empty = torch.exp(ps["d_empty_loc"] + ps["d_empty_scale"] * noise["d_empty"])
    zz = ovae.encode_z(x, sub("enc_z."))
    enc = ovae.encode_nonz(x, chi_ambient.detach(), zz["loc"].detach(), sub("enc_o."), d_empty_loc=ps["d_empty_loc"].detach(),
                           log_count_crossover=cfg["log_count_crossover"], prior_log_cell_counts=cfg["prior_log_cell_counts"],
                           empty_log_count_threshold=cfg["empty_log_count_threshold"], offset=cfg["offset"],
                           training=training, row_keep_scale=row_keep_scale,
                           bn_momentum=0.1 if (update_bn and training) else None)  # nn.BatchNorm1d default momentum
    p_y = enc["p_y"]
    v = p_y + P_LOGIT_SCALE * noise["p_logit_reg"]
    z = zz["loc"] + zz["scale"] * noise["z"]
    prob = p_y.sigmoid().detach()
    d_cell_loc_prior = c(cfg["d_cell_loc_prior"])
    d_cell_loc_gated = prob * enc["d_loc"] + (1 - prob) * d_cell_loc_prior
    d_cell = torch.exp(d_cell_loc_gated + ps["d_cell_scale"] * noise["d_cell"])
    eps_prior = c(cfg["epsilon_prior"])
    eps_gated = prob * enc["epsilon"] + (1 - prob) * 1.0
    epsilon = gamma_rsample(eps_gated * eps_prior, eps_prior, noise["epsilon"])
    log_q1, log_q0 = F.logsigmoid(p_y), F.logsigmoid(-p_y)
    q1, q0 = log_q1.exp(), log_q0.exp()

    # ---------------- model (model.py:232-445), y in {0, 1} enumerated ----------------
    chi = ovae.decode(z, sub("dec."), training=training,
                      bn_momentum=0.01 if (update_bn and training) else None)  # vae/base.py:71 momentum=0.01
    counts = x.sum(-1)
    t: Dict[str, torch.Tensor] = {}
    t["phi"] = Gamma(c(cfg["phi_conc_prior"]), c(cfg["phi_rate_prior"])).log_prob(phi) - Gamma(phi_conc, phi_rate).log_prob(phi)
    t["z"] = (Normal(0.0, 1.0).log_prob(z).sum(-1).sum() - (q1 * Normal(zz["loc"], zz["scale"]).log_prob(z).sum(-1)).sum())
    t["d_cell"] = (LogNormal(d_cell_loc_prior, c(cfg["d_cell_scale_prior"])).log_prob(d_cell)
                   - LogNormal(d_cell_loc_gated, ps["d_cell_scale"]).log_prob(d_cell)).sum()
    if include_rho:
        t["rho"] = (Beta(c(cfg["rho_alpha_prior"]), c(cfg["rho_beta_prior"])).log_prob(rho)
                    - Beta(ps["rho_alpha"], ps["rho_beta"]).log_prob(rho)).sum()
    t["epsilon"] = (Gamma(eps_prior, eps_prior).log_prob(epsilon) - Gamma(eps_gated * eps_prior, eps_prior).log_prob(epsilon)).sum()
    t["d_empty"] = (LogNormal(c(cfg["d_empty_loc_prior"]), c(cfg["d_empty_scale_prior"])).log_prob(d_empty)
                    - LogNormal(ps["d_empty_loc"], ps["d_empty_scale"]).log_prob(d_empty)).sum()
    plp = onb.get_p_logit_prior(counts.log(), d_cell_loc_prior, c(cfg["empty_UMI_threshold"]), c(cfg["p_logit_prior"]))
    t["y"] = (q1 * (F.logsigmoid(plp) - log_q1) + q0 * (F.logsigmoid(-plp) - log_q0)).sum()
    t["p_logit_reg"] = (Normal(c(cfg["p_logit_prior"]), P_LOGIT_SCALE).log_prob(v) - Normal(p_y, P_LOGIT_SCALE).log_prob(v)).sum()

    alpha = phi.reciprocal() + EPS
    if likelihood == "mp":
        like = lambda val, mu, al, lam: _NbpcMp.apply(val, mu, al, lam)
    elif likelihood == "exact":  # NegativeBinomialPoissonConv(max_poisson=100): model.py:349-357 (USE_EXACT_LOG_PROB)
        like = lambda val, mu, al, lam: onb.nbpc_exact_log_prob(val, mu, al, lam, max_poisson=100, float32_tables=False)
    else:
        like = onb.nbpc_approx_log_prob
    obs = torch.zeros((), dtype=dt, device=dev)
    for yv, q in ((1.0, q1), (0.0, q0)):
        y = torch.full((B,), yv, dtype=dt, device=dev)
        mu = onb.calculate_mu(model_type, epsilon=epsilon, d_cell=d_cell, chi=chi, y=y, rho=rho)
        lam = onb.calculate_lambda(model_type, epsilon=epsilon, chi_ambient=chi_ambient, d_empty=d_empty, y=y,
                                   d_cell=d_cell, rho=rho, chi_bar=chi_bar)
        L = like(x, mu + EPS, alpha, lam + EPS).sum(-1)
        obs = obs + (q * L).sum()
    # empty-droplet regulariser (model.py:374-398): epsilon = 1, y = 0, rho and d_cell detached, mask y == 0, scale 0.01
    lam_e = onb.calculate_lambda(model_type, epsilon=torch.ones(B, dtype=dt, device=dev), chi_ambient=chi_ambient, d_empty=d_empty,
                                 y=torch.zeros(B, dtype=dt, device=dev), d_cell=d_cell.detach(), rho=rho.detach(), chi_bar=chi_bar) + EPS
    LE = (torch.xlogy(x, lam_e) - lam_e - torch.lgamma(x + 1)).sum(-1)
    t["obs"] = obs + REG_SCALE_AMBIENT_EXPRESSION * (q0 * LE).sum()

    p_det = p_y.detach()
    crossover = float(np.exp(cfg["log_count_crossover"]))
    empty_m, cell_m = counts < crossover, counts >= crossover
    zero = torch.zeros_like(p_det)
    t["obs_probably_empty_y"] = REG_SCALE_SOFT_SUPERVISION * torch.where(
        empty_m, Normal(-REG_LOGIT_MEAN, REG_LOGIT_SOFT_SCALE).log_prob(p_det), zero).sum()
    t["obs_probably_cell_y"] = REG_SCALE_SOFT_SUPERVISION * torch.where(
        cell_m, Normal(REG_LOGIT_MEAN, REG_LOGIT_SOFT_SCALE).log_prob(p_det), zero).sum()
    surely_cell = counts >= torch.exp(d_cell_loc_prior)
    one = torch.ones((), dtype=dt, device=dev)
    t["epsilon_mean"] = Normal(_masked_median(epsilon, cell_m), 0.01).log_prob(one) if int(surely_cell.sum()) >= 2 \
        else torch.zeros((), dtype=dt, device=dev)
    t["epsilon_empty_mean"] = Normal(_masked_median(epsilon, empty_m), 0.01).log_prob(one)
    t["loss"] = -sum(t.values())
    t["_p_y"] = p_y
    return t
