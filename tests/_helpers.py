"""Shared set-up of the -m gpu parity tests: a small simulated dataset, the product objects built from it exactly as
run_inference builds them, and the product's parameters / buffers / constants re-expressed as inputs of the oracle."""
import numpy as np  # noqa: F401
import torch  # noqa: F401

DEV = "cuda"


def _small_dataset(seed=0, n_genes=96):
    from cellbender_b200 import synth

    sim = synth.simulate_dataset(n_genes=n_genes, cells_of_each_type=[25, 25], n_droplets=600, cell_mean_umi=2000.0,
                                 empty_mean_umi=60.0, dirichlet_alpha=0.5, seed=seed, device="cpu", cell_chunk=64)
    return synth.prepare_dataset(sim, expected_cells=50, total_droplets_included=150)


def _setup(model_type="full", epochs=3, n_genes=96, z_dim=8, hidden=32, exact=False):
    from cellbender_b200 import run

    ds = _small_dataset(n_genes=n_genes)
    args = run.default_args(epochs=epochs, z_dim=z_dim, z_hidden_dims=[hidden], model=model_type, use_exact_log_prob=exact)
    model, opt, svi, train_loader, test_loader = run.setup_inference(ds, args, device=DEV)
    return ds, args, model, opt, svi, train_loader, test_loader


def _oracle_inputs(model, offset):
    params, buffers = {}, {}
    for pre, mod in (("enc_z.", model.encoder["z"]), ("enc_o.", model.encoder["other"]), ("dec.", model.decoder)):
        for k, v in mod.named_parameters():
            params[pre + k] = v.detach().double().cpu().clone().requires_grad_(True)
        for k, v in mod.named_buffers():
            buffers[pre + k] = v.detach().double().cpu().clone()
    for k, v in model.param_store.unconstrained.items():
        params["ps." + k] = v.detach().double().cpu().clone().requires_grad_(True)
    enc_o = model.encoder["other"]
    cfg = dict(chi_bar=model.avg_gene_expression.double().cpu(), log_count_crossover=enc_o.log_count_crossover,
               prior_log_cell_counts=float(enc_o.prior_log_cell_counts),
               empty_log_count_threshold=float(enc_o.empty_log_count_threshold), offset=offset,
               d_cell_loc_prior=float(model.d_cell_loc_prior), d_cell_scale_prior=float(model.d_cell_scale_prior),
               epsilon_prior=float(model.epsilon_prior), phi_conc_prior=float(model.phi_conc_prior),
               phi_rate_prior=float(model.phi_rate_prior), rho_alpha_prior=float(model.rho_alpha_prior),
               rho_beta_prior=float(model.rho_beta_prior), d_empty_loc_prior=float(model.d_empty_loc_prior),
               d_empty_scale_prior=float(model.d_empty_scale_prior), empty_UMI_threshold=float(model.empty_UMI_threshold),
               p_logit_prior=float(model.p_logit_prior))
    return params, buffers, cfg


