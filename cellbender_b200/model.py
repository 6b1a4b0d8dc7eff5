"""The remove-background model: the same object the reference builds (class name, constructor signature,
attributes: reference cellbender/remove_background/model.py:88-230), with the pyro ``model``/``guide`` pair
replaced by an explicit, pyro-free ELBO program whose heavy terms run in the B200 kernels.

Why pyro-free: pyro-ppl is not part of this image (SURVEY.md finding 1) and its per-step Python (two traces,
enumeration, ~40 optimizer objects) is itself slower than the whole B200 step.  The ELBO computed here is the
one ``TraceEnum_ELBO(max_plate_nesting=1)`` computes for ``model``/``guide`` with ``y`` enumerated in the guide
(restated from model.py:232-574); because no reference test pins an ELBO value and pyro
is absent, PARITY OF THE ELBO IS UNPINNED against pyro itself -- it is pinned against the oracle restatement
(oracle/elbo.py), term by term.

Parameter handling mirrors pyro's param store: every ``pyro.param`` site is an UNCONSTRAINED nn.Parameter plus a
``torch.distributions.transform_to(constraint)`` (exp / sigmoid-interval / softmax), which is what pyro
optimises (model.py:245-249, 468-513).
"""
import math
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn
from torch.distributions import constraints, transform_to

from . import consts, gemm, latent, ops
from .exceptions import NanException
from .vae import CompositeEncoder, EncoderInputs


class ParamStore(nn.Module):
    """pyro's param store for this model: unconstrained nn.Parameters + constraint transforms."""

    def __init__(self):
        super().__init__()
        self.unconstrained = nn.ParameterDict()
        self._constraints: Dict[str, Any] = {}

    def setup(self, name: str, init: torch.Tensor, constraint=constraints.real):
        with torch.no_grad():
            u = transform_to(constraint).inv(init.detach().clone().float())
        self.unconstrained[name] = nn.Parameter(u.contiguous())
        self._constraints[name] = constraint

    def keys(self):
        return self.unconstrained.keys()

    def __contains__(self, name):
        return name in self.unconstrained

    def get(self, name: str) -> torch.Tensor:
        return transform_to(self._constraints[name])(self.unconstrained[name])

    def __getitem__(self, name):
        return self.get(name)

    def all(self) -> Dict[str, torch.Tensor]:
        """Every constrained value, computed once (one step reads each several times; the stick-breaking transform of
        chi_ambient alone is ~15 kernels)."""
        return {k: self.get(k) for k in self.unconstrained.keys()}


class _ObsTerm(torch.autograd.Function):
    """sum_b sum_k w[b,k] * L[b,k] with L = (L_y1, L_y0, L_E) of the fused observation kernel.  All gradients
    are produced by the same launch that produces the value, so backward only hands them out (upstream
    gradient must be the scalar 1 that ``loss.backward()`` supplies; it is applied to the small tensors).
    ``w`` = d loss / d L are the enumeration weights -(q1, q0, 0.01 q0); their gradient L carries d/d p_y."""

    @staticmethod
    def forward(ctx, h_dec, w_out, b_out, chi_ambient, coef, alpha, w, model):
        csr, row_ids, bufs = model._obs_ctx
        G = csr.n_genes
        B = h_dec.shape[0]
        logits = bufs["logits"][:B]
        ops.gemm_tf32(h_dec.contiguous(), w_out, B, G, w_out.shape[1], out=logits[:, :G], bias=b_out, split_k=1)
        chi_a = bufs["chi_a"]
        if chi_ambient.data_ptr() != chi_a.data_ptr():
            chi_a[:G].copy_(chi_ambient)
        oo = bufs["obs_out"]
        out = ops.elbo_obs(csr, row_ids, logits, chi_a, bufs["chi_b"], coef.contiguous(), alpha.reshape(1).contiguous(),
                           w.contiguous(), out={"sums": oo["sums"][:B], "dlogits": oo["dlogits"][:B], "qamb": oo["qamb"][:B],
                                                "dchi_ambient": oo["dchi_ambient"], "lse": oo["lse"], "nan_flag": oo["nan_flag"],
                                                "scratch": oo["scratch"]},
                           defer_colsum=model._grad_mode and any(ctx.needs_input_grad),
                           max_poisson=consts.NBPC_EXACT_MAX_POISSON if model.use_exact_log_prob else None)
        ctx.dchi_ready = out["dchi_ready"]
        sums = out["sums"]
        L = sums[:, :3]
        ctx.save_for_backward(h_dec, w_out, sums, w, out["dlogits"], out["dchi_ambient"])
        ctx.G = G
        ctx.alpha_shape = alpha.shape
        ctx.bias_param = b_out  # the nn.Parameter objects (they carry the optimizer's gradient views)
        ctx.w_param = w_out
        neg_obs, obs = ops.obs_loss(sums, w)  # sum_b w . L and its negative (the 'obs' ELBO term)
        ctx.mark_non_differentiable(L, obs)
        ctx.set_materialize_grads(False)  # no zero-filled gradients for L / obs in front of backward (two fill launches)
        return neg_obs, L, obs

    @staticmethod
    def backward(ctx, g, _gL, _gobs):
        if g is None:
            return (None,) * 8
        h_dec, w_out, sums, w, dlogits, dchi_amb = ctx.saved_tensors
        G = ctx.G
        # SVI.step backpropagates from its own persistent scalar 1 (gemm.UNIT_GRAD): the two elementwise products by g
        # below are then identities and their launches (on the critical chain, in front of the decoder's hidden-layer
        # backward) are skipped.  Any other upstream gradient takes the general path.
        unit = gemm.UNIT_GRAD is not None and g.data_ptr() == gemm.UNIT_GRAD.data_ptr()
        B, H = h_dec.shape
        h_c = h_dec.contiguous()
        w_param = ctx.w_param
        d_w_buf, d_w = gemm._grad_target(w_param)
        d_b_buf, d_b = gemm._grad_target(ctx.bias_param)
        main = torch.cuda.current_stream()
        # decoder weight / bias gradients on a side stream, the hidden-activation gradient on this one.
        # CONTRACT: dlogits already is d loss / d logits for an upstream gradient of ONE (the observation kernel folds the
        # loss weights in), and the weight / bias gradients written straight into the optimizer's flat buffer are not
        # rescaled by g -- SVI.step differentiates the loss itself; a caller that backpropagates k * loss must not use the
        # flat-buffer path (d_h, dchi_ambient and the small gradients below do carry g).
        side = gemm.grad_stream(dlogits.device, [w_param])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.gemm_tf32(dlogits, h_c, G, H, B, a_mn=True, b_mn=True, out=d_w_buf, split_k=1)  # dlogits^T . h
            ops.colsum_rows(dlogits, G, out=d_b_buf)
        d_h = ops.gemm_tf32(dlogits, w_out, B, H, G, b_mn=True)              # dlogits[B,G] . Wd[G,H], split-K
        if not unit:
            d_h = d_h * g
        if d_w is None and gemm.early_update_applies(w_param):
            # the decoder weights are no longer read this step: sweep their block now, behind the gradient GEMM
            readers_done = torch.cuda.Event()
            readers_done.record(main)
            side.wait_event(readers_done)
            with torch.cuda.stream(side):
                gemm.FUSED_OPT.early_block_update(w_param)
        if d_w is None and d_b is None:
            h_c.record_stream(side), dlogits.record_stream(side)
            gemm.defer_join(side)  # both gradients went straight to the flat buffer: join at the end of backward
        else:
            main.wait_stream(side)
        d_coef, d_alpha, d_weights = ops.obs_small_grads(sums, w, g.contiguous(), want_dw=ctx.needs_input_grad[6])
        d_alpha = d_alpha.reshape(ctx.alpha_shape)
        if ctx.dchi_ready is not None:
            # the deferred column sums (d loss / d chi_ambient, handed to autograd below) are joined only here, behind the
            # launches above: at the top of this function the wait held the hidden-activation product back until the column
            # sums had finished (r02k timeline: 8 us on the step's critical chain)
            main.wait_event(ctx.dchi_ready)
        return d_h, d_w, d_b, (dchi_amb if unit else dchi_amb * g), d_coef, d_alpha, d_weights, None


class RemoveBackgroundPyroModel(nn.Module):
    """Model + variational guide (reference model.py:88-230 for the constructor; :232-574 for the terms)."""

    def __init__(self, model_type: str, encoder: CompositeEncoder, decoder: nn.Module,
                 dataset_obj_priors: Dict[str, Any], n_analyzed_genes: int, n_droplets: int,
                 analyzed_gene_names: np.ndarray, empty_UMI_threshold: int, log_counts_crossover: float,
                 use_cuda: bool = True, phi_loc_prior: float = consts.PHI_LOC_PRIOR,
                 phi_scale_prior: float = consts.PHI_SCALE_PRIOR, rho_alpha_prior: float = consts.RHO_ALPHA_PRIOR,
                 rho_beta_prior: float = consts.RHO_BETA_PRIOR, epsilon_prior: float = consts.EPSILON_PRIOR,
                 use_exact_log_prob: bool = False, device: Optional[str] = None):
        super().__init__()
        if not use_cuda:
            raise RuntimeError("cellbender_b200 runs on CUDA only (no CPU fallback); pass use_cuda=True")
        if model_type == "simple":
            raise NotImplementedError("the B200 hot path covers model types with empty droplets "
                                      "('ambient', 'swapping', 'full'); 'simple' has no enumerated y")
        self.model_type = model_type
        self.include_empties = True
        self.include_rho = model_type in ("full", "swapping")
        self.n_genes = n_analyzed_genes
        self.n_droplets = n_droplets
        self.analyzed_gene_names = analyzed_gene_names
        self.z_dim = int(decoder.input_dim)
        self.encoder = encoder
        self.encoder_modules = nn.ModuleDict({k: v for k, v in encoder.items()})  # registers the encoder parameters
        self.decoder = decoder
        self.use_exact_log_prob = use_exact_log_prob
        self.loss: Dict[str, Dict[str, list]] = {"train": {"epoch": [], "elbo": []}, "test": {"epoch": [], "elbo": []},
                                                 "learning_rate": {"epoch": [], "value": []}}
        self.log_counts_crossover = log_counts_crossover
        self.counts_crossover = float(np.exp(log_counts_crossover))
        self.device = device or "cuda"
        self.use_cuda = True
        dev = self.device

        p = dataset_obj_priors
        assert p["d_std"] > 0, f"Issue with prior: d_std is {p['d_std']}, but should be > 0."
        assert p["cell_counts"] > 0, f"Issue with prior: cell_counts is {p['cell_counts']}, but should be > 0."
        assert p["empty_counts"] > 0, f"Issue with prior: empty_counts should be > 0, but is {p['empty_counts']}"
        chi_amb0 = torch.as_tensor(p["chi_ambient"]).float()
        chi_bar0 = torch.as_tensor(p["chi_bar"]).float()
        assert np.isclose(float(chi_amb0.sum()), 1.0, atol=1e-5), "Issue with prior: chi_ambient should sum to 1"
        assert np.isclose(float(chi_bar0.sum()), 1.0, atol=1e-5), "Issue with prior: chi_bar should sum to 1"

        def scalar(v):
            return torch.tensor(float(v), dtype=torch.float32, device=dev)

        self.d_cell_loc_prior = scalar(np.log1p(p["cell_counts"]))
        self.d_cell_scale_prior = scalar(p["d_std"])
        self.z_loc_prior = torch.zeros(self.z_dim, device=dev)
        self.z_scale_prior = torch.ones(self.z_dim, device=dev)
        self.epsilon_prior = scalar(epsilon_prior)
        self.phi_loc_prior = scalar(phi_loc_prior)
        self.phi_scale_prior = scalar(phi_scale_prior)
        self.phi_conc_prior = scalar(phi_loc_prior ** 2 / phi_scale_prior ** 2)
        self.phi_rate_prior = scalar(phi_loc_prior / phi_scale_prior ** 2)
        self.d_empty_loc_prior = scalar(np.log1p(p["empty_counts"], dtype=np.float32).item())
        self.d_empty_scale_prior = scalar(p["d_empty_std"])
        self.p_logit_prior = scalar(p["cell_logit"])
        self.chi_ambient_init = chi_amb0.to(dev)
        self.avg_gene_expression = chi_bar0.to(dev)
        self.empty_UMI_threshold = scalar(empty_UMI_threshold)
        self.rho_alpha_prior = scalar(rho_alpha_prior)
        self.rho_beta_prior = scalar(rho_beta_prior)
        # the same priors as python floats: by-value arguments of the fused latent kernels (latent.host_consts)
        self._host = {k: float(getattr(self, k)) for k in (
            "d_cell_loc_prior", "d_cell_scale_prior", "epsilon_prior", "phi_conc_prior", "phi_rate_prior", "rho_alpha_prior",
            "rho_beta_prior", "d_empty_loc_prior", "d_empty_scale_prior", "p_logit_prior", "empty_UMI_threshold")}
        # python-float distribution arguments would be host->device copies inside a captured CUDA graph
        self._k = {"p_logit_scale": scalar(consts.P_LOGIT_SCALE), "reg_logit_mean": scalar(consts.REG_LOGIT_MEAN),
                   "neg_reg_logit_mean": scalar(-consts.REG_LOGIT_MEAN), "reg_logit_soft_scale": scalar(consts.REG_LOGIT_SOFT_SCALE),
                   "eps_median_scale": scalar(0.01), "one": scalar(1.0)}

        # pyro.param sites (model.py:245-249, 468-513)
        ps = ParamStore()
        ps.setup("chi_ambient", self.chi_ambient_init, constraints.simplex)
        ps.setup("d_cell_scale", torch.tensor([consts.D_CELL_SCALE_INIT]), constraints.positive)
        ps.setup("d_empty_loc", self.d_empty_loc_prior, constraints.positive)
        ps.setup("d_empty_scale", self.d_empty_scale_prior, constraints.positive)
        if self.include_rho:
            interval = constraints.interval(consts.RHO_PARAM_MIN, consts.RHO_PARAM_MAX)
            ps.setup("rho_alpha", self.rho_alpha_prior, interval)
            ps.setup("rho_beta", self.rho_beta_prior, interval)
        ps.setup("phi_loc", self.phi_loc_prior, constraints.positive)
        ps.setup("phi_scale", self.phi_scale_prior, constraints.positive)
        self.param_store = ps
        self.to(dev)
        gemm.pad_big_weights(self)  # TMA-addressable rows for the G-wide weights even without an optimizer
        self._d_empty_loc_now = None
        if "other" in encoder:
            encoder["other"].d_empty_loc_fn = lambda: (self._d_empty_loc_now if self._d_empty_loc_now is not None
                                                       else self.param_store["d_empty_loc"])
        self._obs_ctx = None
        self._grad_mode = True
        self._bufs: Dict[int, dict] = {}
        self._retired_bufs: list = []
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)

    # ---- pyro.param(...) compatibility -----------------------------------------------------------------
    def param(self, name: str) -> torch.Tensor:
        return self.param_store[name]

    # ---- buffers reused across steps (fixed addresses: CUDA-graph friendly) -----------------------------
    def _work_buffers(self, B: int) -> dict:
        key = 0
        if key not in self._bufs or self._bufs[key]["cap"] < B:
            G, dev = self.n_genes, self.device
            cap = max(B, 8)
            if key in self._bufs:
                self._retired_bufs.append(self._bufs[key])  # captured step graphs refer to the old buffers by address
            ld = ops.row_pitch(G)
            chi_b = torch.zeros(ld, device=dev)
            chi_b[:G] = self.avg_gene_expression
            self._bufs[key] = {
                "cap": cap, "logits": torch.zeros((cap, ld), device=dev), "chi_a": torch.zeros(ld, device=dev),
                "chi_unit": torch.zeros(ld, device=dev), "chi_b": chi_b,
                "obs_out": {"sums": torch.empty((cap, 12), device=dev), "dlogits": torch.zeros((cap, ld), device=dev),
                            "qamb": torch.empty((cap, ld), device=dev), "dchi_ambient": torch.empty(G, device=dev),
                            "lse": torch.empty(cap, device=dev), "nan_flag": self.nan_flag,
                            "scratch": torch.empty(ops.elbo_obs_scratch_elems(cap), device=dev)}}
        return self._bufs[key]

    # ---- the ELBO --------------------------------------------------------------------------------------
    def ambient(self, B: int = 8):
        """(chi_ambient [G] with autograd history to the unconstrained parameter, chi_ambient / ||chi_ambient||_2):
        ``pyro.param("chi_ambient", constraint=simplex)`` (model.py:245-249) and the unit vector of the encoder's
        ambient-overlap feature (vae/encoder.py:266-270), from one kernel, written into the step's work buffers."""
        bufs = self._work_buffers(B)
        G = self.n_genes
        chi = latent.simplex_param(self.param_store.unconstrained["chi_ambient"], bufs)
        return chi, bufs["chi_unit"][:G]

    PROLOGUE_STREAM = 7  # ops.side_stream index

    def step_prologue(self, B: int, row_keep_scale: Optional[torch.Tensor] = None, csr: Optional[ops.DeviceCSR] = None,
                      row_ids: Optional[torch.Tensor] = None) -> dict:
        """Everything at the head of a step that does not need the gathered minibatch, issued on an auxiliary stream so that
        it runs BESIDE the gather instead of in front of / behind it (r02h / r02j timelines: the one-CTA simplex kernel
        delayed the gather by 12 us, six serial 1.5 us launches sat between the gather and batchnorm0):
          * the Dropout1d row realisation of the encoder (bernoulli + scale) and the constrained ``d_empty_loc``
            -> event ``keep_ready`` (all batchnorm0 needs beside the gathered rows);
          * ``pyro.param("chi_ambient")`` (simplex transform) + the ambient unit vector;
          * with ``csr`` / ``row_ids``: the per-droplet features [B, 4] INCLUDING the dot product with that unit vector
            (vae/encoder.py:266-272) -- the gather itself is then called without the ambient profile; both launches
            reduce the droplet's stored counts in the same order, so the shared features are bit-identical.
        Returns the handle ``elbo_terms(prologue=...)`` joins."""
        dev = self.device
        main = torch.cuda.current_stream()
        draw = row_keep_scale is None and self.encoder["other"].training
        # allocated on the caller's stream (whose later reuse of the blocks is ordered behind the join), filled on the side one
        keep = torch.empty(B, device=dev) if draw else row_keep_scale
        u = self.param_store.unconstrained["d_empty_loc"].detach()
        d_empty_loc = torch.empty_like(u)
        side = ops.side_stream(dev, self.PROLOGUE_STREAM)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if draw:
                keep.bernoulli_(0.5).mul_(2.0)
            torch.exp(u, out=d_empty_loc)  # transform_to(constraints.positive)
            keep_ready = torch.cuda.Event()
            keep_ready.record(side)
            chi_ambient, unit = self.ambient(B)
            feats = ops.gather_feats(csr, row_ids, unit) if csr is not None else None
        return {"row_keep_scale": keep, "d_empty_loc": d_empty_loc, "keep_ready": keep_ready, "chi_ambient": chi_ambient,
                "unit": unit, "feats": feats, "stream": side}

    def elbo_terms(self, csr: ops.DeviceCSR, row_ids: torch.Tensor, inp: EncoderInputs,
                   noise: Optional[Dict[str, torch.Tensor]] = None, row_keep_scale: Optional[torch.Tensor] = None,
                   chi_ambient: Optional[torch.Tensor] = None, prologue: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        """All terms of the ELBO for one minibatch.  Returns a dict of scalar tensors whose
        sum is the ELBO, plus 'loss' = -ELBO.  ``noise`` lets tests inject fixed base randomness
        (keys z, d_cell, d_empty, p_logit_reg: standard normals; phi, epsilon: standard-gamma draws at the current
        concentrations; rho: the Beta draw).  ``prologue``: the handle of ``step_prologue`` (joined here)."""
        B = inp.feats.shape[0]
        self._grad_mode = torch.is_grad_enabled()  # autograd.Function.forward cannot see it (grad mode is off in there)
        main = torch.cuda.current_stream()
        if prologue is not None:
            main.wait_event(prologue["keep_ready"])
            row_keep_scale, d_empty_loc = prologue["row_keep_scale"], prologue["d_empty_loc"]
        else:
            if chi_ambient is None:
                chi_ambient, _ = self.ambient(B)
            d_empty_loc = self.param_store["d_empty_loc"].detach()
        self._d_empty_loc_now = d_empty_loc
        x_in = self.encoder["other"].normalize_input(inp, row_keep_scale)  # batchnorm0 + dropout, on a side stream
        z_loc, z_log_scale = self.encoder["z"].forward_raw(inp)
        if prologue is not None:
            # first consumers of the simplex transform and of the ambient-dot feature: the hand-made encoder features
            main.wait_stream(prologue["stream"])
            chi_ambient = prologue["chi_ambient"]
            if prologue["feats"] is not None:
                inp = EncoderInputs(inp.x_norm, inp.x_lognorm, prologue["feats"], inp.n_genes)
        p_out, eps_out, _ = self.encoder["other"].forward_raw(inp, chi_ambient.detach(), z_loc.detach(),
                                                              row_keep_scale=row_keep_scale, x_in=x_in)
        self._d_empty_loc_now = None
        # guide samples, rate coefficients, enumeration weights and every non-observation term: one kernel
        z, coef, alpha, w, loss_local, terms, enc_out = latent.latent_stage(self, p_out, eps_out, z_loc, z_log_scale,
                                                                            inp.feats, noise)
        t: Dict[str, torch.Tensor] = {name: terms[i] for i, name in enumerate(latent.TERM_NAMES)}
        if not self.include_rho:
            del t["rho"]

        # observation terms: decoder output layer + fused likelihood kernel
        h_dec = self.decoder.hidden(z)
        lin = self.decoder.out_linear
        self._obs_ctx = (csr, row_ids, self._work_buffers(B))
        neg_obs, L, obs = _ObsTerm.apply(h_dec, lin.weight, lin.bias, chi_ambient, coef, alpha, w, self)
        t["obs"] = obs
        self._last_L = L
        t["loss"] = loss_local + neg_obs
        t["_enc"] = {"p_y": enc_out[:, 0], "d_loc": enc_out[:, 1], "epsilon": enc_out[:, 2]}
        return t

    def check_nan(self):
        """Raise NanException like the reference's log_prob does (...ConvApprox.py:129-141); one host sync."""
        if int(self.nan_flag.item()) != 0:
            self.nan_flag.zero_()
            raise NanException(param="mu, alpha, lam (fused likelihood kernel)")
