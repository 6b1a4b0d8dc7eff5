"""Host side of the fused per-droplet latent algebra (csrc/latent.cu): everything of the reference's
``model()`` / ``guide()`` that lives on [B]- or [B, Z]-shaped tensors (reference model.py:232-574), the output shaping
of ``EncodeNonZLatents`` (vae/encoder.py:300-340) and the constraint transforms of the pyro.param sites, as two
autograd nodes whose forward and backward are single one-CTA kernels.

Random numbers stay with torch's CUDA generator (graph-capture safe): one ``randn`` for all Normal base noise and one
``torch._standard_gamma`` for the Gamma / Beta base samples, differentiated with the same implicit reparameterisation
gradients ``Gamma.rsample`` / ``Beta.rsample`` use (``torch._standard_gamma_grad``, ``torch._dirichlet_grad``).
"""
import ctypes
from typing import Dict, Optional

import torch

from . import consts, ops
from ._cabi import current_stream, lib, ptr

TERM_NAMES = ("phi", "z", "d_cell", "rho", "epsilon", "d_empty", "y", "p_logit_reg", "obs_probably_empty_y",
              "obs_probably_cell_y", "epsilon_mean", "epsilon_empty_mean")
SCALAR_PARAMS = ("d_cell_scale", "d_empty_loc", "d_empty_scale", "rho_alpha", "rho_beta", "phi_loc", "phi_scale")
MODEL_TYPE_CODE = {"ambient": 0.0, "swapping": 1.0, "full": 2.0}


def host_consts(model) -> ctypes.Array:
    """consts_h of cb_latent_* (order documented in include/cellbender_b200.h), rebuilt per call because the encoder's
    first-forward offset (vae/encoder.py:300-309) becomes known only after the first minibatch."""
    enc = model.encoder["other"]
    off = 0.0 if enc.offset is None else float(enc.offset["logit_p"])
    vals = [enc.log_count_crossover, enc.prior_log_cell_counts, enc.empty_log_count_threshold, off,
            model._host["d_cell_loc_prior"], model._host["d_cell_scale_prior"], model._host["epsilon_prior"],
            model._host["phi_conc_prior"], model._host["phi_rate_prior"], model._host["rho_alpha_prior"],
            model._host["rho_beta_prior"], model._host["d_empty_loc_prior"], model._host["d_empty_scale_prior"],
            model._host["p_logit_prior"], model._host["empty_UMI_threshold"], model.counts_crossover,
            consts.RHO_PARAM_MIN, consts.RHO_PARAM_MAX, MODEL_TYPE_CODE[model.model_type]]
    n = int(lib().raw("cb_latent_num_consts")())
    assert len(vals) == n, (len(vals), n)
    return (ctypes.c_float * n)(*[float(v) for v in vals])


def _scalar_ptrs(model):
    u = model.param_store.unconstrained
    return [ptr(u[k]) if k in u else None for k in SCALAR_PARAMS]


def _scalar_grad_ptrs(model):
    """Where d loss / d (unconstrained scalar parameter) is written: the optimizer's flat gradient buffer when the
    parameter has been re-homed there (optim.FlatClippedAdam), else a fresh ``.grad``."""
    u = model.param_store.unconstrained
    out = []
    for k in SCALAR_PARAMS:
        if k not in u:
            out.append(None)
            continue
        p = u[k]
        gv = getattr(p, "_cb_grad_view", None)
        if gv is not None:
            p._cb_grad_direct = True
            out.append(ptr(gv))
        else:
            p.grad = torch.zeros_like(p)
            out.append(ptr(p.grad))
    return out


class _LatentStage(torch.autograd.Function):
    """(p_out, eps_out, z_loc, z_log_scale) -> (z, coef, alpha, w, loss_local, terms, enc_out)."""

    @staticmethod
    def forward(ctx, p_out, eps_out, z_loc, z_ls, feats, model, noise: Optional[Dict[str, torch.Tensor]]):
        L = lib()
        st = current_stream()
        dev = feats.device
        B, Z = z_loc.shape
        p_out, eps_out = p_out.contiguous().float(), eps_out.contiguous().float()
        z_loc, z_ls = z_loc.contiguous(), z_ls.contiguous()
        feats = feats.contiguous()
        kh = host_consts(model)
        sp = _scalar_ptrs(model)
        f32 = dict(dtype=torch.float32, device=dev)
        # needs_input_grad reflects requires_grad of the inputs even under torch.no_grad(): without the grad-mode flag the
        # test pass would compute the backward-only reparameterisation gradients on a side stream nobody joins
        need_grad = getattr(model, "_grad_mode", True) and any(ctx.needs_input_grad[:4])

        conc_gamma = torch.empty(1 + 3 * B, **f32)
        conc_beta = torch.empty((B, 2), **f32)
        total_beta = torch.empty((B, 2), **f32)
        L.call("cb_latent_prepare", ptr(feats), ptr(p_out), ptr(eps_out), B, *sp, ctypes.addressof(kh), ptr(conc_gamma),
               ptr(conc_beta), ptr(total_beta), st)
        xx = torch.empty((B, 2), **f32)
        if noise is None:
            normals = torch.randn(B * (Z + 3), **f32)
            n_z, n_row = normals[:B * Z].view(B, Z), normals[B * Z:].view(B, 3)
            gdraw = torch._standard_gamma(conc_gamma)
            L.call("cb_latent_beta_sample", ptr(gdraw), B, ptr(xx), st)
        else:  # injected base noise (tests): standard normals, standard-gamma draws at the current concentrations, rho
            n_z = noise["z"].to(**f32).contiguous()
            n_row = torch.stack((noise["d_cell"], noise["d_empty"], noise["p_logit_reg"]), dim=1).to(**f32).contiguous()
            gdraw = torch.cat((noise["phi"].reshape(1), noise["epsilon"].reshape(B))).to(**f32).contiguous()
            rho = noise["rho"].to(**f32) if "rho" in noise else torch.full((B,), 0.5, **f32)
            xx = torch.stack((rho, 1 - rho), dim=1).contiguous()
        sgg = dgrad = None
        ctx.grads_ready = None
        if need_grad:
            # The implicit reparameterisation gradients are needed only by the backward kernel.  torch's elementwise
            # kernels for them are slow, serial chains (~70 us for 2 x 255 values), so they run on a side stream and
            # overlap the decoder GEMM and the likelihood kernel; backward waits on `grads_ready`.  Their inputs stay
            # alive (saved for backward) and their outputs are first read after that wait, so no allocation crosses
            # streams unsynchronised.
            main = torch.cuda.current_stream()
            side = ops.side_stream(dev, 1)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                sgg = torch._standard_gamma_grad(conc_gamma[:1 + B], gdraw[:1 + B])
                dgrad = torch._dirichlet_grad(xx, conc_beta, total_beta)
                ctx.grads_ready = torch.cuda.Event()
                ctx.grads_ready.record(side)
            sgg.record_stream(main), dgrad.record_stream(main)

        z = torch.empty((B, Z), **f32)
        coef = torch.empty((B, 7), **f32)
        alpha = torch.empty(1, **f32)
        w = torch.empty((B, 3), **f32)
        terms = torch.empty(len(TERM_NAMES), **f32)
        loss_local = torch.empty((), **f32)
        enc_out = torch.empty((B, 4), **f32)
        med_idx = torch.empty(2, dtype=torch.int32, device=dev)
        L.call("cb_latent_forward", ptr(feats), ptr(p_out), ptr(eps_out), ptr(z_loc), ptr(z_ls), z_loc.stride(0), B, Z,
               ptr(n_z), ptr(n_row), ptr(gdraw), ptr(xx), *sp, ctypes.addressof(kh), ptr(z), ptr(coef), ptr(alpha), ptr(w),
               ptr(terms), ptr(loss_local), ptr(enc_out), ptr(med_idx), st)
        if need_grad:
            ctx.save_for_backward(feats, p_out, eps_out, z_loc, z_ls, n_z, n_row, gdraw, xx, sgg, dgrad, med_idx)
            ctx.model = model
            ctx.kh = kh
        ctx.mark_non_differentiable(terms, enc_out)
        # gradients nobody produces (terms, enc_out; any unused output) arrive as None in backward instead of as freshly
        # zero-filled tensors: each of those was a fill launch in front of latent_backward_kernel on the step's critical
        # chain (r02j timeline); the kernel treats a null upstream gradient as zero
        ctx.set_materialize_grads(False)
        return z, coef, alpha, w, loss_local, terms, enc_out

    @staticmethod
    def backward(ctx, d_z, d_coef, d_alpha, d_w, g_local, _terms, _enc):
        feats, p_out, eps_out, z_loc, z_ls, n_z, n_row, gdraw, xx, sgg, dgrad, med_idx = ctx.saved_tensors
        model = ctx.model
        if ctx.grads_ready is not None:
            torch.cuda.current_stream().wait_event(ctx.grads_ready)
        B, Z = z_loc.shape
        f32 = dict(dtype=torch.float32, device=feats.device)
        c = lambda t: None if t is None else t.contiguous()
        d_z, d_coef, d_alpha, d_w, g_local = c(d_z), c(d_coef), c(d_alpha), c(d_w), c(g_local)
        if g_local is None:  # loss_local unused by the caller (the kernel reads a null pointer here as 1)
            g_local = torch.zeros((), **f32)
        d_p_out = torch.empty(B, **f32)
        d_eps_out = torch.empty(B, **f32)
        d_z_loc = torch.empty((B, Z), **f32)
        d_z_ls = torch.empty((B, Z), **f32)
        lib().call("cb_latent_backward", ptr(feats), ptr(p_out), ptr(eps_out), ptr(z_loc), ptr(z_ls), z_loc.stride(0), B, Z,
                   ptr(n_z), ptr(n_row), ptr(gdraw), ptr(xx), ptr(sgg), ptr(dgrad), ptr(med_idx), *_scalar_ptrs(model),
                   ctypes.addressof(ctx.kh), ptr(d_z), ptr(d_coef), ptr(d_alpha), ptr(d_w), ptr(g_local), ptr(d_p_out),
                   ptr(d_eps_out), ptr(d_z_loc), ptr(d_z_ls), d_z_loc.stride(0), *_scalar_grad_ptrs(model), current_stream())
        return d_p_out, d_eps_out, d_z_loc, d_z_ls, None, None, None


def latent_stage(model, p_out, eps_out, z_loc, z_ls, feats, noise=None):
    return _LatentStage.apply(p_out, eps_out, z_loc, z_ls, feats, model, noise)


class _SimplexParam(torch.autograd.Function):
    """transform_to(constraints.simplex)(u) (torch SoftmaxTransform) through cb_simplex_forward / _backward.
    ``bufs`` (a plain dict, invisible to autograd) may hold persistent output buffers 'chi_a' / 'chi_unit'."""

    @staticmethod
    def forward(ctx, u, bufs):
        n = u.numel()
        if bufs is not None:
            y = bufs["chi_a"].detach().narrow(0, 0, n)  # fresh alias: the buffer object itself never gets a grad_fn
            unit = bufs.get("chi_unit")
        else:
            y, unit = torch.empty(n, dtype=torch.float32, device=u.device), None
        lib().call("cb_simplex_forward", ptr(u), n, ptr(y), ptr(unit), current_stream())
        ctx.save_for_backward(y)
        ctx.param = u  # the nn.Parameter object itself (it carries the optimizer's gradient view)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        n = y.numel()
        param = ctx.param
        gv = getattr(param, "_cb_grad_view", None)
        direct = gv is not None and gv.is_contiguous()
        du = gv if direct else torch.empty_like(y)
        lib().call("cb_simplex_backward", n, ptr(y), ptr(gy.contiguous()), ptr(du), current_stream())
        if direct:
            param._cb_grad_direct = True
            return None, None
        return du, None


def simplex_param(u: torch.Tensor, bufs: Optional[dict] = None) -> torch.Tensor:
    """Constrained simplex value softmax(u) [G] of the unconstrained parameter ``u`` [G].  With ``bufs`` the result is
    written into bufs['chi_a'] and bufs['chi_unit'] additionally receives y / ||y||_2 (the ambient unit vector the
    gather kernel dots the counts with)."""
    if not u.is_cuda:
        raise RuntimeError("cellbender_b200 ops require CUDA tensors (no CPU fallback)")
    return _SimplexParam.apply(u, bufs)
