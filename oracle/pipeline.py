"""ORACLE (test infrastructure, NOT the product path): the reference's whole remove-background pipeline restated on
the CPU from the pinned oracle pieces -- training (run_inference / run_training), latents_map, the noise-count
posterior with its Monte-Carlo latent samples, MCKP targets, MCKP estimation and the denoised-count subtraction.
It is the "reference-style pipeline" side of the end-to-end parity check (tools/e2e_parity.py, tests/test_e2e_gpu.py).

Reference locations (relative to /root/reference/cellbender/remove_background):
  run_inference                      run.py:632-856   (seed, batch-size rule, loaders, optimizer, training)
  get_optimizer                      run.py:593-629   (ClippedAdam + OneCycleLR(max_lr = 10 lr, total = n_batches * epochs))
  train_epoch / run_training         train.py:34-72, 108-300
  EncodeNonZLatents first-forward offsets   vae/encoder.py:300-309
  Posterior._get_latents_map         posterior.py:861-917
  Posterior.sample_mu_lambda_alpha   posterior.py:737-782  (guide trace, y := MAP, model replay; the model returns the
                                     lambda of its obs_empty block: epsilon = 1, y = 0 -- model.py:385-393, 445)
  Posterior.noise_log_pdf / _get_cell_noise_count_posterior_coo   posterior.py:424-594, 669-735
  compute_output_denoised_counts_reports_metrics (MCKP targets, estimator, subtraction)   run.py:228-330
  Posterior.compute_denoised_counts  posterior.py:282-330

pyro is absent from this image, so SVI / TraceEnum_ELBO / ClippedAdam are the hand restatements of oracle/elbo.py and
oracle/train_step.py (PARITY UNPINNED against pyro, see those headers); everything downstream of training can be run
with the UNMODIFIED reference functions instead of the oracle ones (``use_reference=True`` in the build container:
``Posterior._log_prob_noise_count_tensor``, ``MultipleChoiceKnapsack``, ``Mean``, ``compute_mean_target_removal_as_function``
imported through oracle/ref_shim.py).
"""
import math
from typing import Callable, Dict, Optional

import numpy as np
import scipy.sparse as sp
import torch
from torch.distributions import transform_to

from . import dataprep as odp
from . import elbo as oelbo
from . import estimation as oest
from . import nbpc as onb
from . import posterior as opost
from . import train_step as ots
from . import vae as ovae

RANDOM_SEED = 1234  # consts.py:4
MAX_BATCH_SIZE = 256  # consts.py:112
CELL_PROB_CUTOFF = 0.5  # consts.py:21


def one_cycle_tables(learning_rate: float, total_steps: int):
    """(lr, beta1) seen by optimizer step k under torch's OneCycleLR(max_lr = 10 lr) around an Adam-family optimizer
    (run.py:604-619; cycle_momentum moves betas[0] 0.95 -> 0.85 -> 0.95)."""
    dummy = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=learning_rate, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.OneCycleLR(dummy, max_lr=learning_rate * 10, total_steps=total_steps)
    lrs, b1s = [], []
    for k in range(total_steps):
        lrs.append(dummy.param_groups[0]["lr"])
        b1s.append(dummy.param_groups[0]["betas"][0])
        dummy.step()
        if k < total_steps - 1:
            sched.step()
    return lrs, b1s


class ScheduledClippedAdam(ots.ClippedAdamPerTensor):
    """pyro ClippedAdam (oracle/train_step.py) with the per-step lr / beta1 of OneCycleLR."""

    def __init__(self, params, lrs, b1s, **kw):
        super().__init__(params, lrs[0], **kw)
        self.lrs, self.b1s = lrs, b1s
        self._b1_prod = 1.0

    @torch.no_grad()
    def step(self):
        k = self.t
        if k >= len(self.lrs):
            raise ValueError(f"Tried to step {k + 1} times. The specified number of total steps is {len(self.lrs)}")
        lr, b1 = self.lrs[k], self.b1s[k]
        self.t += 1
        b2 = self.betas[1]
        # torch.optim.Adam semantics with a time-varying beta1: bias correction uses the CURRENT beta1 ** step
        step_size = lr * math.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        for name, p in self.params.items():
            if p.grad is None:
                continue
            g = p.grad.clamp_(-self.clip, self.clip)
            self.m[name].mul_(b1).add_(g, alpha=1 - b1)
            self.v[name].mul_(b2).addcmul_(g, g, value=1 - b2)
            p.addcdiv_(self.m[name], self.v[name].sqrt().add_(self.eps), value=-step_size)
            p.grad = None


def train(ds, epochs: int, seed: int = RANDOM_SEED, z_dim: int = 64, hidden: int = 512, learning_rate: float = 1e-4,
          training_fraction: float = 0.9, fraction_empties: float = 0.2, model_type: str = "full",
          log: Optional[Callable[[str], None]] = None) -> Dict:
    """run_inference on the host.  Returns dict(params, buffers, cfg, train_loss [epochs], n_batches)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    cm, em = ds.get_count_matrix(), ds.get_count_matrix_empties()
    batch_size = int(min(MAX_BATCH_SIZE, training_fraction * cm.shape[0] / 2))
    # prep_sparse_data_for_training (dataprep.py:196-272): random train / test split of cells and of empties
    tr = np.flatnonzero(np.random.rand(cm.shape[0]) < training_fraction)
    tr_e = np.flatnonzero(np.random.rand(em.shape[0]) < training_fraction)
    loader = odp.HostLoader(cm[tr], em[tr_e], batch_size=batch_size, fraction_empties=fraction_empties)
    n_batches = len(loader)
    params, buffers = ots.init_state(cm.shape[1], z_dim, hidden, ds.priors, seed=seed)
    cfg = ots.make_cfg(ds.priors, ds.empty_UMI_threshold)
    cfg["offset"] = None
    lrs, b1s = one_cycle_tables(learning_rate, n_batches * epochs)
    opt = ScheduledClippedAdam(params, lrs, b1s)
    losses = []
    for epoch in range(1, epochs + 1):
        tot, n_rows = 0.0, 0
        for x in loader:
            B = x.shape[0]
            with torch.no_grad():
                ps = {k[3:]: transform_to(oelbo.CONSTRAINTS[k[3:]])(v) for k, v in params.items() if k.startswith("ps.")}
                conc = {"phi": ps["phi_loc"] ** 2 / ps["phi_scale"] ** 2, "epsilon": torch.full((B,), 50.0),
                        "rho_alpha": ps["rho_alpha"], "rho_beta": ps["rho_beta"]}
            keep = torch.bernoulli(torch.full((B,), 0.5)) * 2.0
            if cfg["offset"] is None:  # first forward fixes the encoder's offsets (encoder.py:300-309)
                cfg["offset"] = first_offsets(x, params, buffers, cfg, keep)
            # epsilon's concentration depends on the encoder output; draw_noise needs it -> one extra no-grad forward
            conc["epsilon"] = epsilon_concentration(x, params, buffers, cfg, keep)
            noise = oelbo.draw_noise(B, z_dim, conc)
            terms = oelbo.elbo(x, params, buffers, cfg, noise, keep, training=True, likelihood="reference",
                               model_type=model_type, update_bn=True)
            terms["loss"].backward()
            opt.step()
            tot += float(terms["loss"].detach())
            n_rows += B
        losses.append(tot / n_rows)
        if log is not None and (epoch == 1 or epoch % 10 == 0 or epoch == epochs):
            log(f"[oracle epoch {epoch:03d}] average training loss {losses[-1]:.4f}")
    return {"params": {k: v.detach() for k, v in params.items()}, "buffers": buffers, "cfg": cfg, "train_loss": losses,
            "n_batches": n_batches, "z_dim": z_dim, "model_type": model_type}


def _split(params, buffers):
    sd = {**params, **buffers}
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    ps = {k[3:]: transform_to(oelbo.CONSTRAINTS[k[3:]])(v) for k, v in params.items() if k.startswith("ps.")}
    return sub, ps


@torch.no_grad()
def _encode(x, params, buffers, cfg, training: bool, keep=None, offset=None):
    sub, ps = _split(params, buffers)
    zz = ovae.encode_z(x, sub("enc_z."))
    enc = ovae.encode_nonz(x, ps["chi_ambient"], zz["loc"], sub("enc_o."), d_empty_loc=ps["d_empty_loc"],
                           log_count_crossover=cfg["log_count_crossover"], prior_log_cell_counts=cfg["prior_log_cell_counts"],
                           empty_log_count_threshold=cfg["empty_log_count_threshold"],
                           offset=offset if offset is not None else cfg["offset"], training=training, row_keep_scale=keep)
    return zz, enc, ps, sub


def first_offsets(x, params, buffers, cfg, keep):
    _, enc, _, _ = _encode(x, params, buffers, cfg, True, keep, offset={"logit_p": 0.0, "epsilon": 0.0})
    return ovae.first_forward_offsets(enc["_p_out"], enc["_eps_out"], enc["_log_sum"], cfg["log_count_crossover"])


def epsilon_concentration(x, params, buffers, cfg, keep):
    """Concentration of the guide's epsilon Gamma for this minibatch (model.py:564-567) -- needed to draw its
    standard-gamma base sample before the differentiable forward."""
    _, enc, _, _ = _encode(x, params, buffers, cfg, True, keep)
    prob = enc["p_y"].sigmoid()
    return (prob * enc["epsilon"] + (1 - prob)) * cfg["epsilon_prior"]


@torch.no_grad()
def latents_map(ds, state, batch: int = 500) -> Dict[str, np.ndarray]:
    """Posterior._get_latents_map (posterior.py:861-917): encoder-only pass over the analysed droplets, eval mode."""
    params, buffers, cfg = state["params"], state["buffers"], state["cfg"]
    cm = ds.get_count_matrix()
    n = cm.shape[0]
    z, d, p, eps = np.zeros((n, state["z_dim"])), np.zeros(n), np.zeros(n), np.zeros(n)
    for a in range(0, n, batch):
        x = torch.from_numpy(np.asarray(cm[a:a + batch].todense(), dtype=np.float32))
        zz, enc, ps, _ = _encode(x, params, buffers, cfg, False)
        sl = slice(a, a + x.shape[0])
        z[sl] = zz["loc"].numpy()
        d[sl] = torch.exp(enc["d_loc"] + ps["d_cell_scale"] ** 2 / 2).numpy()  # LogNormal(loc, scale).mean
        p[sl] = enc["p_y"].sigmoid().numpy()
        eps[sl] = enc["epsilon"].numpy()  # Gamma(eps * k, k).mean
    _, ps = _split(params, buffers)
    return {"z": z, "d": d, "p": p, "epsilon": eps, "phi_loc_scale": [float(ps["phi_loc"]), float(ps["phi_scale"])]}


@torch.no_grad()
def sample_mu_lambda_alpha(x, params, buffers, cfg, model_type: str = "full", device=None):
    """One latent sample of (mu, lam, alpha) for the dense minibatch x as Posterior.sample_mu_lambda_alpha returns it
    (posterior.py:737-782): guide draws, y forced to its MAP, and the model's returned ``lam`` -- which is the lambda of
    the obs_empty regulariser block (epsilon = 1, y = 0; model.py:385-393 overwrites ``lam`` before the return at :445)."""
    from torch.distributions import Beta, Gamma, LogNormal, Normal

    zz, enc, ps, sub = _encode(x, params, buffers, cfg, False)
    B = x.shape[0]
    phi_conc, phi_rate = ps["phi_loc"] ** 2 / ps["phi_scale"] ** 2, ps["phi_loc"] / ps["phi_scale"] ** 2
    phi = Gamma(phi_conc, phi_rate).sample()
    include_rho = model_type in ("full", "swapping")
    rho = Beta(ps["rho_alpha"], ps["rho_beta"]).sample((B,)) if include_rho else None
    d_empty = LogNormal(ps["d_empty_loc"], ps["d_empty_scale"]).sample((B,))
    prob = enc["p_y"].sigmoid()
    z = Normal(zz["loc"], zz["scale"]).sample()
    d_cell = LogNormal(prob * enc["d_loc"] + (1 - prob) * cfg["d_cell_loc_prior"], ps["d_cell_scale"]).sample()
    epsilon = Gamma((prob * enc["epsilon"] + (1 - prob)) * cfg["epsilon_prior"], cfg["epsilon_prior"]).sample()
    y = (enc["p_y"] > 0).float()
    chi = ovae.decode(z, sub("dec."), training=False)
    mu = onb.calculate_mu(model_type, epsilon=epsilon, d_cell=d_cell, chi=chi, y=y, rho=rho)
    dev = x.device
    lam = onb.calculate_lambda(model_type, epsilon=torch.ones(B, device=dev), chi_ambient=ps["chi_ambient"], d_empty=d_empty,
                               y=torch.zeros(B, device=dev), d_cell=d_cell, rho=rho,
                               chi_bar=torch.as_tensor(cfg["chi_bar"]).to(dev))
    return mu, lam, phi.reciprocal()


@torch.no_grad()
def posterior_coo(ds, state, p: np.ndarray, n_samples: int = 20, n_counts_max: int = 20,
                  smallest_log_probability: float = -10.0, posterior_batch_size: int = 128, seed: int = 0,
                  log_prob_fn: Optional[Callable] = None):
    """Posterior._get_cell_noise_count_posterior_coo (posterior.py:424-594).  ``log_prob_fn``: the reference's own
    ``Posterior._log_prob_noise_count_tensor`` when available (build container); default = the pinned oracle restatement.
    Returns (scipy COO [N_all * G_all, n_counts_max], offsets dict)."""
    torch.manual_seed(seed)
    params, buffers, cfg = state["params"], state["buffers"], state["cfg"]
    cm = ds.get_count_matrix()
    cell_logic = p > CELL_PROB_CUTOFF
    if cell_logic.sum() == 0:
        raise RuntimeError("Zero cells found!")
    cell_rows = np.flatnonzero(cell_logic)
    G_all = ds.matrix.shape[1]
    ms, cs, lps, offsets = [], [], [], {}
    SMALLEST_ALLOWED_BATCH = 4
    for a in range(0, cell_rows.size, posterior_batch_size):
        rows = cell_rows[a:a + posterior_batch_size]
        if rows.size < SMALLEST_ALLOWED_BATCH:  # the reference's loader drops such a tail (dataprep.py:155-161)
            break
        x = torch.from_numpy(np.asarray(cm[rows].todense(), dtype=np.float32))
        samples = [sample_mu_lambda_alpha(x, params, buffers, cfg, state["model_type"]) for _ in range(n_samples)]
        if log_prob_fn is None:
            m, c, lp, off = opost.posterior_coo_for_batch(
                x, samples, np.asarray(ds.analyzed_barcode_inds)[rows], np.asarray(ds.analyzed_gene_inds), G_all,
                n_counts_max=n_counts_max, smallest_log_probability=smallest_log_probability)
        else:
            pdf, low = _noise_log_pdf_with(log_prob_fn, x, samples, n_counts_max)
            n_i, g_i, c_i, lpt, offt = opost.sparsify_posterior(pdf, low, x, smallest_log_probability)
            m = opost.m_index(np.asarray(ds.analyzed_barcode_inds)[rows][n_i.numpy()],
                              np.asarray(ds.analyzed_gene_inds)[g_i.numpy()], G_all)
            c, lp = c_i.numpy().astype(np.int32), lpt.numpy()
            offn = offt.numpy().astype(np.int64)
            off = {int(k): int(v) for k, v in zip(m[offn != 0], offn[offn != 0])}
        ms.append(m), cs.append(c), lps.append(lp), offsets.update(off)
    m, c, lp = np.concatenate(ms), np.concatenate(cs), np.concatenate(lps)
    shape = (int(np.prod(ds.matrix.shape, dtype=np.uint64)), n_counts_max)
    return sp.coo_matrix((lp, (m, c)), shape=shape), offsets


def _noise_log_pdf_with(log_prob_fn, data, samples, n_counts_max):
    """noise_log_pdf (posterior.py:698-735) around a supplied ``_log_prob_noise_count_tensor``."""
    total = low = None
    for s, (mu, lam, alpha) in enumerate(samples, start=1):
        lp, low = log_prob_fn(data=data, mu_est=mu + 1e-30, lambda_est=lam + 1e-30, alpha_est=alpha + 1e-30,
                              n_counts_max=n_counts_max, debug=False)
        lp = lp - torch.logsumexp(lp, dim=-1, keepdim=True)
        total = lp if s == 1 else torch.logaddexp(total + torch.log(torch.tensor(1.0 - 1.0 / s)),
                                                  lp + torch.log(torch.tensor(1.0 / s)))
    return total, low


def denoise(ds, coo: sp.coo_matrix, offsets: Dict[int, int], p: np.ndarray, fpr: float = 0.01, reference_estimation=None):
    """MCKP targets at ``fpr`` (run.py:264-282, posterior.py:1589-1650), MCKP estimation (estimation.py:420-702) and the
    subtraction of the noise from the called cells (posterior.py:315-328).  ``reference_estimation``: module
    ``cellbender.remove_background.estimation`` + ``posterior`` of the unmodified reference (tuple) when available.
    Returns dict(noise_csr, denoised_csc, targets)."""
    shape = ds.matrix.shape
    cell_inds = np.asarray(ds.analyzed_barcode_inds)[p > CELL_PROB_CUTOFF]
    keep = np.zeros(shape[0], dtype=bool)
    keep[cell_inds] = True
    cell_counts = sp.csr_matrix(sp.diags(keep.astype(ds.matrix.dtype)) @ ds.matrix)
    if reference_estimation is None:
        per_cell = oest.mean_target_removal_per_gene(coo, offsets, shape, cell_counts, n_cells=len(cell_inds), fpr=fpr)
        targets = per_cell * len(cell_inds)
        noise_csr = oest.estimate_mckp(coo, offsets, shape, targets)
    else:
        ref_est, ref_post = reference_estimation
        conv = ref_post.IndexConverter(total_n_cells=shape[0], total_n_genes=shape[1])
        fun = ref_post.compute_mean_target_removal_as_function(
            noise_count_posterior_coo=coo, noise_offsets=offsets, index_converter=conv, raw_count_csr_for_cells=cell_counts,
            n_cells=len(cell_inds), device="cpu", per_gene=True)
        targets = (fun(fpr) * len(cell_inds)).detach().cpu().numpy()
        noise_csr = ref_est.MultipleChoiceKnapsack(index_converter=conv).estimate_noise(
            noise_log_prob_coo=coo, noise_offsets=offsets, noise_targets_per_gene=targets, device="cpu",
            use_multiple_processes=False)
    denoised = (cell_counts - noise_csr).tocsc()
    return {"noise_csr": noise_csr, "denoised_csc": denoised, "targets": np.asarray(targets), "cell_inds": cell_inds}
