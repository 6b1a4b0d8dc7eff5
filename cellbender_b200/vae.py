"""Encoder / decoder networks with the reference's module interface (same class names, constructor
arguments, attribute paths and state_dict keys -- they are part of the checkpoint layout, SURVEY.md
Appendix B), computing through the B200 kernels.

Reference: cellbender/remove_background/vae/{base,encoder,decoder}.py.

What runs where
  * the three G -> 512 input layers and the 512 -> G output layer (the only large contractions) go through
    ``gemm.linear_big`` (hand-written tcgen05 GEMM when enabled, see csrc/gemm_tf32.cu);
  * the dense minibatch never exists as raw counts: the loader hands over ``EncoderInputs`` produced by the
    gather kernel (x/(sum+eps), log1p of it, per-droplet features);
  * the decoder stops at the logits: softmax is fused into the likelihood kernel (``Decoder.forward`` still
    returns chi for callers that want it).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import consts
from . import gemm, ops
from .gemm import bn0_dropout, bn_act, linear_big, multi_linear_big


# run the two towers of EncodeNonZLatents on separate streams (see forward_raw)
TOWER_STREAMS = True


@dataclass
class EncoderInputs:
    """Output of the gather kernel for one minibatch (cb_csr_gather_rows)."""

    x_norm: torch.Tensor  # [B, ld]  x / (sum + 1e-5)            (EncodeZ input, transform 'normalize')
    x_lognorm: torch.Tensor  # [B, ld]  log1p(x / (sum + 1e-5))  (EncodeNonZLatents input, 'log_normalize')
    feats: torch.Tensor  # [B, 4]   sum, nnz, dot(x, ambient unit vector), max
    n_genes: int

    @property
    def counts(self) -> torch.Tensor:
        return self.feats[:, 0]


class FullyConnectedLayer(nn.Module):
    """Linear (+ BatchNorm1d(momentum=0.01, eps=0.001)) (+ activation), reference vae/base.py:31-82."""

    def __init__(self, input_dim: int, output_dim: int, activation=nn.ReLU, use_batch_norm: bool = False,
                 use_layer_norm: bool = False, dropout_rate: Optional[float] = None):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        modules: List[nn.Module] = []
        if dropout_rate is not None:
            modules.append(nn.Dropout(p=dropout_rate))
        modules.append(nn.Linear(in_features=input_dim, out_features=output_dim))
        if use_batch_norm:
            modules.append(nn.BatchNorm1d(num_features=output_dim, momentum=0.01, eps=0.001))
        if use_layer_norm:
            modules.append(nn.LayerNorm(normalized_shape=output_dim, elementwise_affine=False))
        if activation is not None:
            modules.append(activation() if isinstance(activation, type) else activation)
        self.layer = nn.Sequential(*modules)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layer(x)


class FullyConnectedNetwork(nn.Module):
    """Stack of FullyConnectedLayer, reference vae/base.py:85-148."""

    def __init__(self, input_dim: int, hidden_dims: List[int], output_dim: int, hidden_activation=nn.ReLU(),
                 output_activation=None, use_batch_norm: bool = False, use_layer_norm: bool = False,
                 norm_output: bool = False, dropout_rate: Optional[float] = None, dropout_input: bool = False):
        super().__init__()
        if use_layer_norm and use_batch_norm:
            raise UserWarning("You are trying to use both batch norm and layer norm. That's probably too much norm.")
        dims = list(zip([input_dim] + hidden_dims, hidden_dims + [output_dim]))
        n_layers = 1 + len(hidden_dims)
        self.network = nn.Sequential(*[
            FullyConnectedLayer(
                input_dim=i, output_dim=j,
                activation=hidden_activation if (layer < n_layers - 1) else output_activation,
                use_batch_norm=use_batch_norm if ((layer < n_layers - 1) or norm_output) else False,
                use_layer_norm=use_layer_norm if ((layer < n_layers - 1) or norm_output) else False,
                dropout_rate=None if ((layer == 0) and not dropout_input) else dropout_rate)
            for layer, (i, j) in enumerate(dims)])

    def forward(self, x: torch.Tensor):
        return self.network(x)


def transform_input(x: torch.Tensor, transform: Optional[str], eps: float = 1e-5) -> torch.Tensor:
    """Dense-tensor version of the reference's transform_input (vae/encoder.py:343-381)."""
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
    raise NotImplementedError("Specified an input transform that is not supported.  Choose from 'log' or 'normalize'.")


def _as_inputs(x, n_genes: int, chi_ambient=None, d_empty_loc=None) -> EncoderInputs:
    """Accept either EncoderInputs (hot path) or a dense count tensor (reference call signature)."""
    if isinstance(x, EncoderInputs):
        return x
    x = x.reshape(-1, n_genes).float()
    xn = transform_input(x, "normalize")
    if chi_ambient is not None and d_empty_loc is not None:
        amb = d_empty_loc.exp().detach() * chi_ambient.detach()
        dot = (x * (amb / torch.linalg.vector_norm(amb))).sum(-1)
    else:
        dot = torch.zeros(x.shape[0], device=x.device)
    feats = torch.stack([x.sum(-1), (x > 0).sum(-1).float(), dot, x.max(-1).values], dim=1)
    return EncoderInputs(xn, xn.log1p(), feats, n_genes)


class EncodeZ(FullyConnectedNetwork):
    """Gene expression -> latent z (loc, scale).  Reference vae/encoder.py:55-125."""

    def __init__(self, input_dim: int, hidden_dims: List[int], output_dim: int, input_transform: Optional[str] = None,
                 **kwargs):
        assert len(hidden_dims) > 0, "EncodeZ needs to have at least one hidden layer"
        super().__init__(input_dim=input_dim, hidden_dims=hidden_dims[:-1], output_dim=hidden_dims[-1],
                         hidden_activation=nn.Softplus(), output_activation=nn.Softplus(), norm_output=True, **kwargs)
        self.transform = input_transform
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.loc_out = nn.Linear(hidden_dims[-1], output_dim)
        self.sig_out = nn.Linear(hidden_dims[-1], output_dim)

    def forward_raw(self, inp: EncoderInputs):
        """(loc, log_scale) [B, Z]: the two output Linears before ``exp`` (what the fused latent kernel consumes)."""
        assert self.transform == "normalize", "the B200 path implements input_transform='normalize' (run.py:707)"
        first = self.network[0].layer
        lin = first[0]
        h = linear_big(inp.x_norm, lin, n_in=self.input_dim)
        for mod in list(first)[1:]:
            h = mod(h)
        for layer in list(self.network)[1:]:
            h = layer(h)
        # both heads read the same hidden activations: one node, one paired input-gradient GEMM
        return multi_linear_big(h, [(self.loc_out, None), (self.sig_out, None)], n_in=h.shape[1])

    def forward(self, x, **kwargs) -> Dict[str, torch.Tensor]:
        loc, log_scale = self.forward_raw(_as_inputs(x, self.input_dim))
        return {"loc": loc.squeeze(), "scale": torch.exp(log_scale).squeeze()}


class EncodeNonZLatents(nn.Module):
    """Gene expression -> (logit cell probability, cell size, epsilon).  Reference vae/encoder.py:132-340.
    ``d_empty_loc_fn`` replaces the reference's ``pyro.param("d_empty_loc")`` lookups (encoder.py:262,268,330):
    it is set by the model object to return the current parameter value."""

    def __init__(self, n_genes: int, z_dim: int, log_count_crossover: float, prior_log_cell_counts: float,
                 empty_log_count_threshold: float, prior_logit_cell_prob: float, input_transform: Optional[str] = None):
        super().__init__()
        self.n_genes = n_genes
        self.z_dim = z_dim
        self.transform = input_transform
        self.output_dim = 1
        self.INITIAL_WEIGHT_FOR_LOG_COUNTS = 1.0
        self.P_OUTPUT_SCALE = 5.0
        self.log_count_crossover = log_count_crossover
        self.empty_log_count_threshold = empty_log_count_threshold
        self.prior_log_cell_counts = prior_log_cell_counts
        self.prior_logit_cell_prob = prior_logit_cell_prob
        self.EPS_OUTPUT_SCALE = 0.2
        self.EPS_OUTPUT_MEAN = 0.6931
        additional_features_p = 4
        self.layer1 = nn.Linear(additional_features_p + self.n_genes, 512)
        self.batchnorm1 = nn.BatchNorm1d(num_features=512)
        self.layer2 = nn.Linear(additional_features_p + 512, 512)
        self.batchnorm2 = nn.BatchNorm1d(num_features=512)
        self.layer3 = nn.Linear(additional_features_p + 512, 1)
        additional_features_eps = 4
        self.eps_network = nn.Sequential(
            nn.Linear(additional_features_eps + self.n_genes, 512), nn.BatchNorm1d(num_features=512), nn.Softplus(),
            nn.Linear(512, 512), nn.BatchNorm1d(num_features=512), nn.Softplus(), nn.Linear(512, 1))
        self.softplus = nn.Softplus()
        self.dropout50 = nn.Dropout1d(p=0.5)
        self.offset: Optional[dict] = None
        self.x_scaling = None
        self.batchnorm0 = nn.BatchNorm1d(num_features=self.n_genes)
        self.d_empty_loc_fn = None  # callable -> 0-dim tensor (set by RemoveBackgroundPyroModel)
        # storage hint (gemm.padded_layout): the G-wide block of these weights starts 4 feature columns into the row; 28
        # floats of lead padding put it on a 128-byte line (names, shapes, values and state_dict keys are unaffected)
        self.layer1.weight._cb_lead = ops.ROW_ALIGN - additional_features_p
        self.eps_network[0].weight._cb_lead = ops.ROW_ALIGN - additional_features_eps

    def _d_empty_loc(self, device):
        if self.d_empty_loc_fn is None:
            raise RuntimeError("EncodeNonZLatents.d_empty_loc_fn is not set (the model object sets it)")
        return self.d_empty_loc_fn().detach().to(device)

    def normalize_input(self, inp: "EncoderInputs", row_keep_scale: Optional[torch.Tensor] = None):
        """dropout50(batchnorm0(x)) (encoder.py:239, 279) issued on an auxiliary stream: it depends on the counts only, so
        it overlaps EncodeZ (and, in backward, its gradient kernel overlaps EncodeZ's weight-gradient GEMM).  Returns a
        handle for ``forward_raw(x_in=...)``, which joins the stream."""
        B = inp.feats.shape[0]
        if self.training and row_keep_scale is None:
            row_keep_scale = torch.empty(B, device=inp.feats.device).bernoulli_(0.5).mul_(2.0)
        if not (inp.x_lognorm.is_cuda and TOWER_STREAMS):
            return bn0_dropout(inp.x_lognorm, self.batchnorm0, row_keep_scale, self.training, self.n_genes), None
        main, side = torch.cuda.current_stream(), ops.side_stream(inp.feats.device, 3)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            x_in = bn0_dropout(inp.x_lognorm, self.batchnorm0, row_keep_scale, self.training, self.n_genes)
        if row_keep_scale is not None:
            row_keep_scale.record_stream(side)
        return x_in, side

    def forward_raw(self, x, chi_ambient: Optional[torch.Tensor], z: torch.Tensor,
                    row_keep_scale: Optional[torch.Tensor] = None, x_in=None):
        """Raw outputs of the two towers, ``layer3(...)`` and ``eps_network(...)`` [B], before the offset / gating
        algebra of encoder.py:300-340 (which the fused latent kernel applies on the training path).  Returns
        (p_out, eps_out, inputs)."""
        assert self.transform == "log_normalize", "the B200 path implements input_transform='log_normalize' (run.py:717)"
        G = self.n_genes
        d_empty_loc = self._d_empty_loc(z.device) if chi_ambient is not None else None
        inp = _as_inputs(x, G, chi_ambient, d_empty_loc)
        # BatchNorm over genes + Dropout1d(0.5), which on a 2-D input drops whole droplet rows (torch 2.11):
        # one fused kernel (csrc/bn_act.cu) writing a 16-byte-row buffer the tensor-core GEMMs can read through TMA.
        B = inp.feats.shape[0]
        if x_in is None:
            x_in = self.normalize_input(inp, row_keep_scale)
        x_in, bn0_stream = x_in
        main = torch.cuda.current_stream()
        e = self.eps_network[0]
        eps_bn1, eps_lin2, eps_bn2, eps_lin3 = self.eps_network[1], self.eps_network[3], self.eps_network[4], self.eps_network[6]
        fork = x_in.is_cuda and TOWER_STREAMS and bn0_stream is not None
        lazy_p, lazy_e = gemm.LazyFeat(), gemm.LazyFeat()
        if fork:
            # The two G-wide products Linear(cat(feat, x_in)) of the towers need only batchnorm0's output for their gene
            # block: they are issued right behind it, each on its own stream, while THIS stream computes the hand-made
            # features (which need z, i.e. the tail of EncodeZ).  The feature columns join in bn_act (gemm.LazyFeat).
            side_p, side_e = bn0_stream, ops.side_stream(x_in.device, 2)
            bn0_done = torch.cuda.Event()
            bn0_done.record(bn0_stream)
            side_e.wait_event(bn0_done)
            with torch.cuda.stream(bn0_stream):  # makes the node's "current stream" the one x_in lives on
                h1, he = multi_linear_big(x_in, [(self.layer1, lazy_p), (e, lazy_e)], n_in=G, col_offset=4,
                                          bias_grad_elsewhere=True, streams=[None, side_e])
        # hand-made features (encoder.py:250-287): log1p(counts), log1p(nnz), overlap with the empty-droplet size /
        # ambient dot product, ||z||_2 -- one kernel instead of ~15 elementwise launches
        p_feat, e_feat = ops.encoder_features(inp.feats, z, d_empty_loc)
        log_sum = p_feat[:, 0:1]
        lazy_p.value, lazy_e.value = p_feat, e_feat
        if not fork:
            if bn0_stream is not None:
                main.wait_stream(bn0_stream)
                x_in.record_stream(main)
            h1, he = multi_linear_big(x_in, [(self.layer1, lazy_p), (e, lazy_e)], n_in=G, col_offset=4, bias_grad_elsewhere=True)

        # The two towers are independent chains of small, latency-bound launches ([B, 512] tensors): the epsilon tower
        # runs on an auxiliary stream, the cell-probability tower on the calling one (parallel branches of the captured
        # step graph; autograd replays each backward node on its forward stream).
        def eps_tower():
            h = bn_act(he, eps_bn1, "softplus", self.training, bias_of=e, feat=e_feat)
            h = linear_big(h, eps_lin2, n_in=h.shape[1], bias_grad_elsewhere=True)
            h = bn_act(h, eps_bn2, "softplus", self.training, bias_of=eps_lin2)
            return ops.head(h, eps_lin3)

        if fork:
            feats_done = torch.cuda.Event()
            feats_done.record(main)
            side_e.wait_event(feats_done)
            with torch.cuda.stream(side_e):
                eps_out = eps_tower()
            main.wait_stream(side_p)  # layer1's product
            h1.record_stream(main), x_in.record_stream(main)
        # hidden layers: Linear (tensor-core GEMM + rank-4 feature update) -> fused BatchNorm + Softplus kernel
        x_ = bn_act(h1, self.batchnorm1, "softplus", self.training, bias_of=self.layer1, feat=p_feat)
        h2 = linear_big(x_, self.layer2, n_in=x_.shape[1], col_offset=4, feat=p_feat, bias_grad_elsewhere=True)
        x_ = bn_act(h2, self.batchnorm2, "softplus", self.training, bias_of=self.layer2)
        p_out = ops.head(x_, self.layer3, p_feat)
        if fork:
            main.wait_stream(side_e)
            he.record_stream(side_e), e_feat.record_stream(side_e), eps_out.record_stream(main)
        else:
            eps_out = eps_tower()

        if self.offset is None:  # data-dependent module state fixed by the first minibatch (encoder.py:300-309)
            ls = log_sum.squeeze(-1)
            self.offset = dict()
            cells = ls > self.log_count_crossover
            assert cells.sum() > 4, "Fewer than 4 cells passed to encoder minibatch"
            self.offset["logit_p"] = p_out.mean().item()
            self.offset["epsilon"] = eps_out[cells].mean().item()
        return p_out, eps_out, inp

    def forward(self, x, chi_ambient: Optional[torch.Tensor], z: torch.Tensor, row_keep_scale: Optional[torch.Tensor] = None,
                **kwargs) -> Dict[str, torch.Tensor]:
        p_out, eps_out, inp = self.forward_raw(x, chi_ambient, z, row_keep_scale)
        counts = inp.feats[:, 0:1]
        ls = counts.log1p().squeeze(-1)
        dlt = ls - self.log_count_crossover
        p_y_logit = (p_out - self.offset["logit_p"]) + dlt.abs().pow(0.5) * torch.sign(dlt) * 10.0
        beta = 50.0
        alpha = (beta * (ls - self.prior_log_cell_counts)).sigmoid()
        p_y_logit = (1.0 - alpha) * p_y_logit + alpha * consts.REG_LOGIT_MEAN
        alpha_empty = (beta * (ls - self.empty_log_count_threshold)).sigmoid()
        p_y_logit = alpha_empty * p_y_logit + (1.0 - alpha_empty) * (-1 * consts.REG_LOGIT_MEAN)

        epsilon = 2.0 * (eps_out * self.EPS_OUTPUT_SCALE - 1.0986122886681098).sigmoid() + 0.5
        d_empty = self._d_empty_loc(z.device).exp() if chi_ambient is not None else torch.zeros((), device=z.device)
        d_loc = (self.softplus((self.softplus(counts.squeeze(-1) / (epsilon + 1e-2) - d_empty) + 1e-10).log()
                               - self.log_count_crossover) + self.log_count_crossover)
        return {"p_y": p_y_logit, "d_loc": d_loc, "epsilon": epsilon}


class CompositeEncoder(dict):
    """Several encoders run on the same input (reference vae/encoder.py:11-52)."""

    def __init__(self, module_dict):
        super().__init__(module_dict)
        self.module_dict = module_dict

    def __call__(self, **kwargs):
        return self.forward(**kwargs)

    def forward(self, **kwargs) -> Dict[str, torch.Tensor]:
        out = dict()
        out["z"] = self.module_dict["z"].forward(**kwargs)
        for key, value in self.module_dict.items():
            if key == "z":
                continue
            out[key] = value.forward(**kwargs, z=out["z"]["loc"].detach())
            if key == "other":
                for subkey, v in out[key].items():
                    out[subkey] = v
                del out[key]
        return out


class Decoder(FullyConnectedNetwork):
    """Latent z -> fractional expression chi (softmax).  Reference vae/decoder.py:6-47."""

    def __init__(self, input_dim: int, **kwargs):
        super().__init__(input_dim=input_dim, **kwargs)
        self.input_dim = input_dim
        self.softmax = nn.Softmax(dim=-1)

    def hidden(self, z: torch.Tensor) -> torch.Tensor:
        """Everything before the final 512 -> G Linear."""
        h = z
        for layer in list(self.network)[:-1]:
            mods = list(layer.layer)
            if (len(mods) == 3 and isinstance(mods[0], nn.Linear) and isinstance(mods[1], nn.BatchNorm1d)
                    and isinstance(mods[2], nn.ReLU) and h.is_cuda):
                # Linear -> BatchNorm1d -> ReLU (vae/base.py:66-79): tensor-core GEMM + one fused kernel
                h = linear_big(h.contiguous(), mods[0], n_in=mods[0].in_features, bias_grad_elsewhere=True)
                h = bn_act(h, mods[1], "relu", self.training, bias_of=mods[0])
            else:
                h = layer(h)
        return h

    @property
    def out_linear(self) -> nn.Linear:
        return self.network[-1].layer[0]

    def logits(self, z: torch.Tensor) -> torch.Tensor:
        lin = self.out_linear
        return torch.nn.functional.linear(self.hidden(z), lin.weight, lin.bias)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        return self.softmax(self.logits(z))
