"""ORACLE (test infrastructure, NOT the product path): functional CPU restatement of the
reference's encoder / decoder networks, operating on a plain ``state_dict`` whose keys are
the reference's own attribute paths (they are part of the checkpoint layout, SURVEY.md
Appendix B).  dtype-generic plain torch; differentiable.

Reference locations (relative to /root/reference/cellbender/remove_background/vae):
  transform_input          encoder.py:343-381
  EncodeZ.forward          encoder.py:112-125   (Linear G->H + Softplus; loc/sig heads)
  EncodeNonZLatents.forward encoder.py:241-340
  Decoder.forward          decoder.py:46-47, base.py:66-79 (Linear+BN(momentum .01, eps 1e-3)+ReLU, Linear, softmax)
  CompositeEncoder.forward  encoder.py:33-52
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

REG_LOGIT_MEAN = 10.0  # consts.py:84
LOG3 = 1.0986122886681098  # encoder.py:327


def transform_input(x, transform: Optional[str], eps: float = 1e-5):
    if transform is None:
        return x
    if transform == "log":
        return x.log1p()
    if transform == "normalize":
        return x / (x.sum(dim=-1, keepdim=True) + eps)
    if transform == "normalize_log":
        x = x.log1p()
        return x / (x.sum(dim=-1, keepdim=True) + eps)
    if transform == "log_normalize":
        return (x / (x.sum(dim=-1, keepdim=True) + eps)).log1p()
    raise NotImplementedError(transform)


def batch_norm(x, sd: Dict[str, torch.Tensor], prefix: str, training: bool, eps: float = 1e-5,
               momentum: Optional[float] = None):
    """nn.BatchNorm1d forward: batch statistics (biased variance) in training, running
    statistics in eval.  ``momentum`` (training only): also update the running statistics in
    place as torch does (running = (1 - m) running + m batch, UNBIASED batch variance); None
    leaves them alone (they do not affect the forward value of the same step)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = x.mean(dim=0)
        var = x.var(dim=0, unbiased=False)
        if momentum is not None:
            with torch.no_grad():
                n = x.shape[0]
                sd[prefix + ".running_mean"].mul_(1 - momentum).add_(momentum * mean.detach())
                sd[prefix + ".running_var"].mul_(1 - momentum).add_(momentum * var.detach() * (n / max(n - 1, 1)))
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var + eps) * w + b


def encode_z(x, sd: Dict[str, torch.Tensor], prefix: str = ""):
    """EncodeZ with a single hidden layer and input_transform='normalize' (run.py:701-708)."""
    x_ = transform_input(x, "normalize")
    h = F.softplus(F.linear(x_, sd[prefix + "network.0.layer.0.weight"], sd[prefix + "network.0.layer.0.bias"]))
    loc = F.linear(h, sd[prefix + "loc_out.weight"], sd[prefix + "loc_out.bias"])
    scale = torch.exp(F.linear(h, sd[prefix + "sig_out.weight"], sd[prefix + "sig_out.bias"]))
    return {"loc": loc, "scale": scale}


def encode_nonz(x, chi_ambient, z_loc, sd: Dict[str, torch.Tensor], *, d_empty_loc, log_count_crossover: float,
                prior_log_cell_counts: float, empty_log_count_threshold: float, offset: Dict[str, float],
                training: bool, row_keep_scale: Optional[torch.Tensor] = None, prefix: str = "",
                bn_momentum: Optional[float] = None):
    """EncodeNonZLatents.forward (encoder.py:241-340) with input_transform='log_normalize'.

    ``row_keep_scale`` [B] is the Dropout1d(0.5) realisation for 2-D input: 0 for a
    dropped row, 2 for a kept one (torch 2.11 probe); None = no dropout.
    ``offset`` holds the first-forward heuristics {'logit_p', 'epsilon'} (encoder.py:300-309);
    use :func:`first_forward_offsets` to compute them."""
    counts = x.sum(dim=-1, keepdim=True)
    log_sum = counts.log1p()
    log_nnz = (x > 0).sum(dim=-1, keepdim=True).to(x.dtype).log1p()
    d_empty_loc = torch.as_tensor(d_empty_loc, dtype=x.dtype, device=x.device)
    overlap = -0.5 * (torch.clamp(log_sum - d_empty_loc, min=0.0) / 0.1).pow(2)
    x_amb = d_empty_loc.exp() * chi_ambient.unsqueeze(0)
    x_amb_norm = x_amb / torch.linalg.vector_norm(x_amb, ord=2, dim=-1, keepdim=True)
    eps_overlap = (x_amb_norm * x).sum(dim=-1, keepdim=True)

    xt = transform_input(x, "log_normalize")
    x_in = batch_norm(xt, sd, prefix + "batchnorm0", training, momentum=bn_momentum)
    if training and row_keep_scale is not None:
        x_in = x_in * row_keep_scale.unsqueeze(-1)
    znorm = torch.linalg.vector_norm(z_loc, ord=2, dim=-1, keepdim=True)
    p_feat = torch.cat((log_sum, log_nnz, overlap, znorm), dim=-1)
    e_feat = torch.cat((log_sum, log_nnz, eps_overlap, znorm), dim=-1)

    def lin(name, inp):
        return F.linear(inp, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])

    h = F.softplus(batch_norm(lin("layer1", torch.cat((p_feat, x_in), -1)), sd, prefix + "batchnorm1", training, momentum=bn_momentum))
    h = F.softplus(batch_norm(lin("layer2", torch.cat((p_feat, h), -1)), sd, prefix + "batchnorm2", training, momentum=bn_momentum))
    p_out = lin("layer3", torch.cat((p_feat, h), -1)).squeeze(-1)

    e = F.softplus(batch_norm(lin("eps_network.0", torch.cat((e_feat, x_in), -1)), sd, prefix + "eps_network.1", training, momentum=bn_momentum))
    e = F.softplus(batch_norm(lin("eps_network.3", e), sd, prefix + "eps_network.4", training, momentum=bn_momentum))
    eps_out = lin("eps_network.6", e).squeeze(-1)

    ls = log_sum.squeeze(-1)
    dlt = ls - log_count_crossover
    p_y = (p_out - offset["logit_p"]) + dlt.abs().pow(0.5) * torch.sign(dlt) * 10.0
    beta = 50.0
    a = (beta * (ls - prior_log_cell_counts)).sigmoid()
    p_y = (1.0 - a) * p_y + a * REG_LOGIT_MEAN
    a_e = (beta * (ls - empty_log_count_threshold)).sigmoid()
    p_y = a_e * p_y + (1.0 - a_e) * (-REG_LOGIT_MEAN)

    epsilon = 2.0 * (eps_out * 0.2 - LOG3).sigmoid() + 0.5
    d_empty = d_empty_loc.exp()
    d_loc = F.softplus((F.softplus(counts.squeeze(-1) / (epsilon + 1e-2) - d_empty) + 1e-10).log()
                       - log_count_crossover) + log_count_crossover
    return {"p_y": p_y, "d_loc": d_loc, "epsilon": epsilon, "_p_out": p_out, "_eps_out": eps_out, "_log_sum": ls}


def first_forward_offsets(p_out, eps_out, log_sum, log_count_crossover: float) -> Dict[str, float]:
    """encoder.py:300-309: logit_p offset = mean(p_out); epsilon offset = mean(eps_out[cells])
    (the epsilon offset is stored but never used afterwards, as in the reference)."""
    cells = log_sum > log_count_crossover
    assert int(cells.sum()) > 4, "Fewer than 4 cells passed to encoder minibatch"
    return {"logit_p": float(p_out.mean()), "epsilon": float(eps_out[cells].mean())}


def decode(z, sd: Dict[str, torch.Tensor], training: bool, prefix: str = "", return_logits: bool = False,
           bn_momentum: Optional[float] = None):
    """Decoder with hidden_dims=[H]: Linear(Z->H)+BatchNorm1d(H, eps=1e-3)+ReLU, Linear(H->G), softmax."""
    h = F.linear(z, sd[prefix + "network.0.layer.0.weight"], sd[prefix + "network.0.layer.0.bias"])
    h = F.relu(batch_norm(h, sd, prefix + "network.0.layer.1", training, eps=1e-3, momentum=bn_momentum))
    logits = F.linear(h, sd[prefix + "network.1.layer.0.weight"], sd[prefix + "network.1.layer.0.bias"])
    return logits if return_logits else torch.softmax(logits, dim=-1)
