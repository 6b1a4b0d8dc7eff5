"""Drop-in adapter: makes an importable ``cellbender`` installation (the reference) run its hot path through this
package.  The reference has no plugin / FFI interface; its boundary is four Python call sites, and
this is the binding a maintainer would add there:

    import cellbender_b200.adapter as adapter
    adapter.install()          # before cellbender.remove_background.run is used

| reference call site (cellbender/remove_background/...)                         | replaced by |
|---|---|
| ``data/dataprep.py:153-193, 275-292``  ``DataLoader.__next__`` + ``sparse_collate`` | ``ResidentDataLoader``: same constructor, RNG consumption and iteration protocol; the dense fp32 minibatch the reference's ``svi.step(x)`` expects comes from the device-resident CSR through ``cb_csr_gather_rows`` (no scipy slicing, no 34 MB H2D copy) |
| ``distributions/NegativeBinomialPoissonConvApprox.py:105-143``  ``log_prob``        | ``ops.nbpc_approx_log_prob`` (``cb_nbpc_approx_log_prob_fwd / _bwd``; differentiable in mu, alpha, lam; raises the reference's ``NanException``) |
| ``estimation.py:42-52``  ``EstimationMethod.estimate_noise`` of MAP / Mean / ThresholdCDF / SingleSample / MultipleChoiceKnapsack | the same-named classes of ``cellbender_b200.estimation`` (``cb_segment_*``, ``cb_mckp_*``); same kwargs, same scipy CSR result |
| ``posterior.py:424-594, 669-859``  ``Posterior._get_cell_noise_count_posterior_coo`` (which drives ``noise_log_pdf`` / ``sample_mu_lambda_alpha`` / ``_log_prob_noise_count_tensor``) | ``posterior_from_reference(...)`` builds this package's ``Posterior`` from the reference objects (``state_dict``s of the encoders / decoder, pyro param store, encoder offsets) and returns its COO + offsets |

The per-minibatch ELBO itself (``model()`` / ``guide()`` under ``TraceEnum_ELBO``) is pyro's tracing machinery; running
training through the fused step means calling ``cellbender_b200.run.run_inference`` instead of the reference's
``run_inference`` -- same arguments object, same returned model attributes (``loss``, encoders / decoder with the
reference's ``state_dict`` keys), see INTEGRATION.md.

Nothing here is imported by the package itself; ``install()`` fails loudly when the reference cannot be imported.
"""
from typing import Dict

import numpy as np
import torch

from . import dataprep, estimation, ops, vae
from .exceptions import NanException
from .model import RemoveBackgroundPyroModel
from .posterior import Posterior
from .vae import CompositeEncoder

_INSTALLED: list = []  # (owner object, attribute name, original value)


class ResidentDataLoader(dataprep.DataLoader):
    """The reference's ``DataLoader`` protocol (iteration yields the dense fp32 ``[B, G]`` minibatch tensor on the device)
    on top of the device-resident loader: host-side index logic and numpy RNG consumption are those of the reference,
    the dense tensor comes from the gather kernel."""

    def __init__(self, dataset, empty_drop_dataset, batch_size=128, fraction_empties=0.2, shuffle=True, sort_by=None,
                 use_cuda=True):
        super().__init__(dataset, empty_drop_dataset, batch_size=batch_size, fraction_empties=fraction_empties, shuffle=shuffle,
                         sort_by=sort_by, use_cuda=use_cuda)

    def __next__(self) -> torch.Tensor:
        batch = super().__next__()
        return None if batch is None else batch.dense()


def _nbpc_approx_log_prob(self, value):
    """NegativeBinomialPoissonConvApprox.log_prob(value) through cb_nbpc_approx_log_prob_fwd (NanException included)."""
    return ops.nbpc_approx_log_prob(value, self.mu, self.alpha, self.lam)


def product_model_from_reference(ref_model, ref_param_store: Dict[str, torch.Tensor], dataset_obj, device="cuda"):
    """A ``cellbender_b200`` model object carrying the trained state of a reference ``RemoveBackgroundPyroModel``:
    encoder / decoder ``state_dict``s load with ``strict=True`` (the attribute paths are the checkpoint layout,
    SURVEY.md Appendix B), ``pyro.param`` values (constrained) are written through the constraint transforms, the
    encoder's first-forward offsets are copied."""
    enc_z_ref, enc_o_ref, dec_ref = ref_model.encoder["z"], ref_model.encoder["other"], ref_model.decoder
    n_genes = int(enc_o_ref.n_genes)
    hidden = [enc_z_ref.loc_out.in_features]
    enc_z = vae.EncodeZ(input_dim=n_genes, hidden_dims=hidden, output_dim=enc_z_ref.loc_out.out_features, use_batch_norm=False,
                        use_layer_norm=False, input_transform="normalize")
    enc_o = vae.EncodeNonZLatents(n_genes=n_genes, z_dim=enc_o_ref.z_dim, log_count_crossover=enc_o_ref.log_count_crossover,
                                  prior_log_cell_counts=enc_o_ref.prior_log_cell_counts,
                                  empty_log_count_threshold=enc_o_ref.empty_log_count_threshold,
                                  prior_logit_cell_prob=enc_o_ref.prior_logit_cell_prob, input_transform="log_normalize")
    dec = vae.Decoder(input_dim=enc_z_ref.loc_out.out_features, hidden_dims=hidden[::-1], use_batch_norm=True,
                      use_layer_norm=False, output_dim=n_genes)
    enc_z.load_state_dict(enc_z_ref.state_dict(), strict=True)
    enc_o.load_state_dict(enc_o_ref.state_dict(), strict=True)
    dec.load_state_dict(dec_ref.state_dict(), strict=True)
    enc_o.offset = None if enc_o_ref.offset is None else dict(enc_o_ref.offset)
    model = RemoveBackgroundPyroModel(
        model_type=ref_model.model_type, encoder=CompositeEncoder({"z": enc_z, "other": enc_o}), decoder=dec,
        dataset_obj_priors=dataset_obj.priors, n_analyzed_genes=n_genes, n_droplets=int(ref_model.n_droplets),
        analyzed_gene_names=getattr(ref_model, "analyzed_gene_names", np.arange(n_genes)),
        empty_UMI_threshold=float(ref_model.empty_UMI_threshold), log_counts_crossover=float(ref_model.log_counts_crossover),
        use_cuda=True, device=device)
    from torch.distributions import transform_to

    with torch.no_grad():
        for name, u in model.param_store.unconstrained.items():
            if name in ref_param_store:
                value = torch.as_tensor(ref_param_store[name]).detach().float().to(u.device).reshape(u.shape)
                u.copy_(transform_to(model.param_store._constraints[name]).inv(value))
    model.loss = getattr(ref_model, "loss", model.loss)
    return model.eval()


def posterior_from_reference(ref_posterior, ref_param_store: Dict[str, torch.Tensor], device="cuda") -> Posterior:
    """This package's ``Posterior`` for the dataset / trained model of a reference ``Posterior`` object."""
    model = product_model_from_reference(ref_posterior.vi_model, ref_param_store, ref_posterior.dataset_obj, device=device)
    return Posterior(_DatasetView(ref_posterior.dataset_obj), model, posterior_batch_size=ref_posterior.posterior_batch_size,
                     debug=ref_posterior.debug)


class _DatasetView:
    """What ``cellbender_b200.Posterior`` reads from a dataset object, served from the reference's
    ``SingleCellRNACountsDataset`` (its count matrix lives in ``data['matrix']``)."""

    def __init__(self, ref_ds):
        self._ds = ref_ds
        self.analyzed_gene_inds = ref_ds.analyzed_gene_inds
        self.analyzed_barcode_inds = ref_ds.analyzed_barcode_inds
        self.matrix = ref_ds.data["matrix"]
        self.priors = ref_ds.priors

    def get_count_matrix(self):
        return self._ds.get_count_matrix()

    def restore_eliminated_features_in_cells(self, *a, **k):
        return self._ds.restore_eliminated_features_in_cells(*a, **k)


def _get_cell_noise_count_posterior_coo(self, n_samples: int = 20, y_map: bool = True, n_counts_max: int = 20,
                                        smallest_log_probability: float = -10.0):
    """Replacement of the reference method of the same name (posterior.py:424-594)."""
    import pyro

    store = {k: pyro.param(k).detach() for k in pyro.get_param_store().keys()}
    post = posterior_from_reference(self, store, device="cuda")
    post._latents = self.latents_map  # cell calls exactly as the reference made them
    coo = post._get_cell_noise_count_posterior_coo(n_samples=n_samples, y_map=y_map, n_counts_max=n_counts_max,
                                                   smallest_log_probability=smallest_log_probability)
    self._noise_count_posterior_coo = coo
    self._noise_count_posterior_coo_offsets = post._noise_count_posterior_coo_offsets
    return coo


def install() -> list:
    """Patch the four call sites inside the importable reference.  Returns [(owner, attribute, original value)]."""
    if _INSTALLED:
        return _INSTALLED
    try:
        import cellbender.remove_background.data.dataprep as ref_dp
        import cellbender.remove_background.distributions.NegativeBinomialPoissonConvApprox as ref_nb
        import cellbender.remove_background.estimation as ref_est
        import cellbender.remove_background.posterior as ref_post
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise RuntimeError("cellbender_b200.adapter.install(): the reference package `cellbender` is not importable") from exc

    def patch(owner, name, new):
        _INSTALLED.append((owner, name, owner.__dict__[name] if name in getattr(owner, "__dict__", {}) else getattr(owner, name)))
        setattr(owner, name, new)

    patch(ref_dp, "DataLoader", ResidentDataLoader)
    patch(ref_nb.NegativeBinomialPoissonConvApprox, "log_prob", _nbpc_approx_log_prob)
    for cls in ("MAP", "Mean", "ThresholdCDF", "SingleSample", "MultipleChoiceKnapsack"):
        patch(getattr(ref_est, cls), "estimate_noise", getattr(estimation, cls).estimate_noise)
    patch(ref_post.Posterior, "_get_cell_noise_count_posterior_coo", _get_cell_noise_count_posterior_coo)
    # the exception type the reference's training loop catches (train.py:277)
    import cellbender.remove_background.exceptions as ref_exc

    if ref_exc.NanException is not NanException:
        patch(ops, "NanException", ref_exc.NanException)
    return _INSTALLED


def uninstall():
    """Undo ``install()`` (tests)."""
    while _INSTALLED:
        owner, name, orig = _INSTALLED.pop()
        setattr(owner, name, orig)
