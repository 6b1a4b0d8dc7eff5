"""ORACLE (test infrastructure / timed CPU baseline, NOT the product path): one reference-style training step on
the host -- the reference's loader (oracle/dataprep.py), the dense eager ELBO in float32 with the reference's own
likelihood formula (oracle/elbo.py, likelihood='reference'), autograd backward, and one pyro-ClippedAdam update per
parameter tensor (pyro.optim.clipped_adam.ClippedAdam.step restated: clamp +-10, Adam moments, denominator
sqrt(v) + eps, step lr * sqrt(1 - b2^t) / (1 - b1^t)).  It stands in for ``svi.step`` of the reference
(train.py:56-60) WITHOUT pyro's tracing overhead, i.e. it is a lower bound on the reference's CPU step time.
"""
import math
import time
from typing import Dict

import numpy as np
import torch

from . import elbo as oelbo


def init_state(n_genes: int, z_dim: int, hidden: int, priors: Dict, seed: int = 1234, dtype=torch.float32, device="cpu"):
    """Random-init parameters of the reference architecture (torch default inits), as plain tensors keyed like the
    reference's state_dicts."""
    torch.manual_seed(seed)
    G, H, Z = n_genes, hidden, z_dim
    lin = lambda i, o: torch.nn.Linear(i, o)
    params, buffers = {}, {}

    def add_lin(name, i, o):
        l = lin(i, o)
        params[name + ".weight"], params[name + ".bias"] = l.weight.detach().to(dtype), l.bias.detach().to(dtype)

    def add_bn(name, n):
        params[name + ".weight"], params[name + ".bias"] = torch.ones(n, dtype=dtype), torch.zeros(n, dtype=dtype)
        buffers[name + ".running_mean"], buffers[name + ".running_var"] = torch.zeros(n, dtype=dtype), torch.ones(n, dtype=dtype)

    add_lin("enc_z.network.0.layer.0", G, H), add_lin("enc_z.loc_out", H, Z), add_lin("enc_z.sig_out", H, Z)
    add_bn("enc_o.batchnorm0", G)
    add_lin("enc_o.layer1", G + 4, 512), add_bn("enc_o.batchnorm1", 512), add_lin("enc_o.layer2", 516, 512)
    add_bn("enc_o.batchnorm2", 512), add_lin("enc_o.layer3", 516, 1)
    add_lin("enc_o.eps_network.0", G + 4, 512), add_bn("enc_o.eps_network.1", 512), add_lin("enc_o.eps_network.3", 512, 512)
    add_bn("enc_o.eps_network.4", 512), add_lin("enc_o.eps_network.6", 512, 1)
    add_lin("dec.network.0.layer.0", Z, H), add_bn("dec.network.0.layer.1", H), add_lin("dec.network.1.layer.0", H, G)
    from torch.distributions import transform_to

    init = {"chi_ambient": torch.as_tensor(priors["chi_ambient"], dtype=dtype), "d_cell_scale": torch.tensor([0.02], dtype=dtype),
            "d_empty_loc": torch.tensor(float(np.log1p(priors["empty_counts"])), dtype=dtype),
            "d_empty_scale": torch.tensor(float(priors["d_empty_std"]), dtype=dtype),
            "rho_alpha": torch.tensor(1.5, dtype=dtype), "rho_beta": torch.tensor(50.0, dtype=dtype),
            "phi_loc": torch.tensor(0.2, dtype=dtype), "phi_scale": torch.tensor(0.2, dtype=dtype)}
    for k, v in init.items():
        params["ps." + k] = transform_to(oelbo.CONSTRAINTS[k]).inv(v)
    params = {k: v.detach().to(device).requires_grad_(True) for k, v in params.items()}
    buffers = {k: v.to(device) for k, v in buffers.items()}
    return params, buffers


def make_cfg(priors: Dict, empty_UMI_threshold: float, offset=None, device="cpu"):
    return dict(chi_bar=torch.as_tensor(priors["chi_bar"]).to(device), log_count_crossover=float(priors["log_counts_crossover"]),
                prior_log_cell_counts=float(np.log1p(priors["cell_counts"])),
                empty_log_count_threshold=float(np.log1p(empty_UMI_threshold)), offset=offset or {"logit_p": 0.0, "epsilon": 0.0},
                d_cell_loc_prior=float(np.log1p(priors["cell_counts"])), d_cell_scale_prior=float(priors["d_std"]),
                epsilon_prior=50.0, phi_conc_prior=1.0, phi_rate_prior=5.0, rho_alpha_prior=1.5, rho_beta_prior=50.0,
                d_empty_loc_prior=float(np.log1p(priors["empty_counts"])), d_empty_scale_prior=float(priors["d_empty_std"]),
                empty_UMI_threshold=float(empty_UMI_threshold), p_logit_prior=float(priors["cell_logit"]))


class ClippedAdamPerTensor:
    def __init__(self, params: Dict[str, torch.Tensor], lr: float, clip=10.0, betas=(0.9, 0.999), eps=1e-8):
        self.params, self.lr, self.clip, self.betas, self.eps, self.t = params, lr, clip, betas, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        step_size = self.lr * math.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        for k, p in self.params.items():
            if p.grad is None:
                continue
            g = p.grad.clamp_(-self.clip, self.clip)
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            p.addcdiv_(self.m[k], self.v[k].sqrt().add_(self.eps), value=-step_size)
            p.grad = None


def reference_step(x: torch.Tensor, params, buffers, cfg, optimizer: ClippedAdamPerTensor, z_dim: int) -> float:
    """One reference-style training step on the dense minibatch x [B, G] (float32), on x's device (the host for the
    CPU baseline; a GPU for the "reference torch-CUDA path" stand-in, where x arrives by a per-step H2D copy exactly as
    the reference's loader delivers it, dataprep.py:291-292)."""
    B, dev = x.shape[0], x.device
    with torch.no_grad():
        ps = {k[3:]: torch.distributions.transform_to(oelbo.CONSTRAINTS[k[3:]])(v) for k, v in params.items() if k.startswith("ps.")}
        conc = {"phi": ps["phi_loc"] ** 2 / ps["phi_scale"] ** 2, "epsilon": torch.full((B,), 50.0, device=dev),
                "rho_alpha": ps["rho_alpha"], "rho_beta": ps["rho_beta"]}
    noise = oelbo.draw_noise(B, z_dim, conc, device=dev)
    keep = torch.bernoulli(torch.full((B,), 0.5, device=dev)) * 2.0
    terms = oelbo.elbo(x, params, buffers, cfg, noise, keep, training=True, likelihood="reference")
    terms["loss"].backward()
    optimizer.step()
    return float(terms["loss"].detach())
