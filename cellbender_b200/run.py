"""Construction of the inference objects exactly as the reference's run_inference does it (reference
cellbender/remove_background/run.py:632-856): seeds, EncodeZ / EncodeNonZLatents / Decoder hyper-parameters,
batch size rule, train/test split, optimizer.  The CLI, file I/O, report and checkpoint tarball around it are out of
scope -- a caller that has a prepared dataset object gets the same model trained the same way.
"""
import argparse
import os
import logging
import random
from typing import Optional, Tuple

import numpy as np
import torch

from . import consts
from .dataprep import DataLoader, prep_sparse_data_for_training
from .model import RemoveBackgroundPyroModel
from .optim import FlatClippedAdam, get_optimizer
from .train import SVI, run_training
from .vae import CompositeEncoder, Decoder, EncodeNonZLatents, EncodeZ

logger = logging.getLogger("cellbender")


def default_args(**overrides) -> argparse.Namespace:
    """The reference's argparse defaults on the hot path (argparser.py:119-182, 234-243, 295-304, 335-344)."""
    d = dict(model="full", epochs=150, z_dim=64, z_hidden_dims=[512], learning_rate=1e-4, training_fraction=0.9,
             fraction_empties=0.2, constant_learning_rate=False, epoch_elbo_fail_fraction=None,
             final_elbo_fail_fraction=None, posterior_batch_size=consts.PROB_POSTERIOR_BATCH_SIZE, fpr=[0.01],
             estimator="mckp", use_cuda=True, debug=False, checkpoint_min=7.0,
             use_cuda_graph=True)
    d.update(overrides)
    return argparse.Namespace(**d)


def set_rng_seed(seed: int):
    """pyro.set_rng_seed (run.py:664-666): python, numpy and torch (all devices)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def build_model(dataset_obj, args, device: Optional[str] = None) -> RemoveBackgroundPyroModel:
    """run.py:695-743."""
    n_genes = int(dataset_obj.analyzed_gene_inds.size)
    encoder_z = EncodeZ(input_dim=n_genes, hidden_dims=args.z_hidden_dims, output_dim=args.z_dim, use_batch_norm=False,
                        use_layer_norm=False, input_transform="normalize")
    encoder_other = EncodeNonZLatents(
        n_genes=n_genes, z_dim=args.z_dim, log_count_crossover=dataset_obj.priors["log_counts_crossover"],
        prior_log_cell_counts=np.log1p(dataset_obj.priors["cell_counts"]),
        empty_log_count_threshold=np.log1p(dataset_obj.empty_UMI_threshold),
        prior_logit_cell_prob=dataset_obj.priors["cell_logit"], input_transform="log_normalize")
    encoder = CompositeEncoder({"z": encoder_z, "other": encoder_other})
    decoder = Decoder(input_dim=args.z_dim, hidden_dims=args.z_hidden_dims[::-1], use_batch_norm=True,
                      use_layer_norm=False, output_dim=n_genes)
    gene_names = np.array([f"g{n:05d}" for n in dataset_obj.analyzed_gene_inds])
    return RemoveBackgroundPyroModel(
        model_type=args.model, encoder=encoder, decoder=decoder, dataset_obj_priors=dataset_obj.priors,
        n_analyzed_genes=n_genes, n_droplets=int(dataset_obj.analyzed_barcode_inds.size),
        analyzed_gene_names=gene_names, empty_UMI_threshold=dataset_obj.empty_UMI_threshold,
        log_counts_crossover=dataset_obj.priors["log_counts_crossover"], use_cuda=True, device=device,
        # consts.USE_EXACT_LOG_PROB of the reference (consts.py:68, a module constant there): the exact convolution likelihood
        use_exact_log_prob=bool(getattr(args, "use_exact_log_prob", consts.USE_EXACT_LOG_PROB)))


def setup_inference(dataset_obj, args, device: Optional[str] = None, total_epochs_for_testing_only: Optional[int] = None,
                    rank: int = 0, world_size: int = 1, dp_mode: Optional[str] = None):
    """Everything run_inference builds before training starts (run.py:658-786).  With world_size > 1 every rank
    builds the identical model (same seed) and takes a disjoint shard of the training droplets (see dp.py);
    ``dp_mode``: None (fused NVLink exchange when peer memory can be mapped, else NCCL), 'p2p' or 'sharded'."""
    # the reference switches distribution argument validation off (run.py:659-660); it costs a host sync per check
    torch.distributions.Distribution.set_default_validate_args(False)
    set_rng_seed(consts.RANDOM_SEED)
    count_matrix = dataset_obj.get_count_matrix()
    model = build_model(dataset_obj, args, device=device)
    frac = args.training_fraction
    batch_size = int(min(consts.MAX_BATCH_SIZE, frac * dataset_obj.analyzed_barcode_inds.size / 2))
    empties = dataset_obj.get_count_matrix_empties()
    if world_size > 1:
        from .dp import shard_rows

        count_matrix = count_matrix[shard_rows(count_matrix.shape[0], rank, world_size)]
        np.random.seed(consts.RANDOM_SEED + 7919 * rank)  # per-rank loader stream (empties choice, shuffles)
    train_loader, test_loader = prep_sparse_data_for_training(
        dataset=count_matrix, empty_drop_dataset=empties, batch_size=batch_size, training_fraction=frac,
        fraction_empties=args.fraction_empties, shuffle=True, device=device)
    n_batches = len(train_loader)
    symmetric_group = None
    if world_size > 1:
        from .dp import agree_on_steps, p2p_group

        # every rank must take the same number of steps per epoch (and sweep with the same OneCycleLR table): the shards
        # and the random train / test split make the loaders' lengths differ by one between ranks
        n_batches = agree_on_steps(n_batches, device=device)
        train_loader.limit_batches(n_batches)
        symmetric_group = p2p_group() if dp_mode in (None, "p2p") else None
    opt_args = dict(n_batches=n_batches, batch_size=train_loader.batch_size, epochs=args.epochs,
                    learning_rate=args.learning_rate, constant_learning_rate=args.constant_learning_rate,
                    total_epochs_for_testing_only=total_epochs_for_testing_only)
    if symmetric_group is not None:
        # the parameters are re-homed by the optimizer, so probe the symmetric-memory rendezvous BEFORE building it: if
        # this system cannot map peer memory (every rank sees the same failure), NCCL carries the exchange (dp.attach)
        try:
            import torch.distributed._symmetric_memory as symm_mem

            probe = symm_mem.empty(64, dtype=torch.float32, device=device or "cuda")
            symm_mem.rendezvous(probe, symmetric_group)
        except Exception as exc:  # noqa: BLE001
            if dp_mode == "p2p":
                raise
            logger.warning(f"symmetric memory unavailable ({type(exc).__name__}: {exc}); data-parallel exchange through NCCL")
            symmetric_group = None
    scheduler = get_optimizer(model, symmetric_group=symmetric_group, **opt_args)
    svi = SVI(model, scheduler, use_cuda_graph=getattr(args, "use_cuda_graph", True))
    return model, scheduler, svi, train_loader, test_loader


def run_inference(dataset_obj, args, device: Optional[str] = None, total_epochs_for_testing_only: Optional[int] = None,
                  rank: int = 0, world_size: int = 1) -> Tuple[RemoveBackgroundPyroModel, FlatClippedAdam, DataLoader, DataLoader]:
    """Full inference procedure (run.py:632-856 without checkpoint tarballs and the ELBO-retry re-entry).  With
    ``world_size`` > 1 (one process per GPU, torch.distributed initialised) training is data parallel, see dp.py."""
    model, scheduler, svi, train_loader, test_loader = setup_inference(
        dataset_obj, args, device=device, total_epochs_for_testing_only=total_epochs_for_testing_only, rank=rank,
        world_size=world_size)
    if world_size > 1:
        from . import dp

        dp.attach(svi, world_size)
    if args.epochs > 0:
        logger.info("Running inference...")
        run_training(model=model, svi=svi, train_loader=train_loader, test_loader=test_loader, epochs=args.epochs,
                     test_freq=5, epoch_elbo_fail_fraction=args.epoch_elbo_fail_fraction,
                     final_elbo_fail_fraction=args.final_elbo_fail_fraction, learning_rate=args.learning_rate)
        logger.info("Inference procedure complete.")
    return model, scheduler, train_loader, test_loader


def compute_output_denoised_counts(posterior, args, fpr_list=None):
    """The estimation half of ``compute_output_denoised_counts_reports_metrics`` (run.py:228-330) without the file
    writers / report: estimator choice (:253-286), MCKP per-gene targets from the Mean estimator at each FPR
    (posterior.py:1589-1650), ``Posterior.compute_denoised_counts`` and ``restore_eliminated_features_in_cells``.
    Returns {fpr: scipy CSC [all barcodes, all genes]}."""
    from .estimation import (MAP, Mean, MultipleChoiceKnapsack, SingleSample, ThresholdCDF,
                             compute_mean_target_removal_as_function)
    from .sparse_utils import combine_csr

    posterior.cell_noise_count_posterior_coo()
    ds = posterior.dataset_obj
    noise_target_fun = None
    estimators = {"map": MAP, "mean": Mean, "sample": SingleSample, "cdf": ThresholdCDF, "mckp": MultipleChoiceKnapsack}
    if args.estimator not in estimators:
        raise ValueError('Input --estimator must be one of ["map", "mean", "sample", "cdf", "mckp"]')
    estimator = estimators[args.estimator]
    if args.estimator == "mckp":
        logger.info("Computing target noise counts per gene for MCKP estimator")
        count_matrix = ds.matrix  # all barcodes
        cell_inds = np.asarray(ds.analyzed_barcode_inds)[posterior.latents_map["p"] > consts.CELL_PROB_CUTOFF]
        empty_inds = np.setdiff1d(np.arange(count_matrix.shape[0]), cell_inds)
        # csr_set_rows_to_zero(count_matrix, empty_inds) of run.py:268, NOT in place: the reference's helper zeroes a
        # csr_matrix argument in place (its own data["matrix"] is CSC, so it works on a converted copy); the raw counts
        # of this dataset object must survive for restore_eliminated_features_in_cells
        cell_counts = combine_csr(count_matrix, zero_rows=empty_inds)
        # the device-resident COO (same entries as the scipy one, already (m, c)-sorted) feeds the estimators directly
        coo = posterior._device_posterior if posterior._device_posterior is not None else posterior._noise_count_posterior_coo
        per_cell = compute_mean_target_removal_as_function(
            noise_count_posterior_coo=coo, noise_offsets=posterior._noise_count_posterior_coo_offsets,
            index_converter=posterior.index_converter, raw_count_csr_for_cells=cell_counts, n_cells=len(cell_inds),
            device=posterior.device, per_gene=True)
        noise_target_fun = lambda x: per_cell(x) * len(cell_inds)
    out = {}
    for fpr in (fpr_list if fpr_list is not None else args.fpr):
        noise_targets = noise_target_fun(fpr).detach().cpu().numpy() if noise_target_fun is not None else None
        denoised = posterior.compute_denoised_counts(
            estimator_constructor=estimator, noise_targets_per_gene=noise_targets, q=getattr(args, "cdf_threshold_q", 0.5),
            alpha=getattr(args, "prq_alpha", 1.0), device=posterior.device, use_multiple_processes=False)
        denoised = ds.restore_eliminated_features_in_cells(denoised, posterior.latents_map["p"])
        assert np.all(denoised.data >= 0), "Negative count matrix entries in output"
        out[fpr] = denoised
    return out


def run_remove_background(dataset_obj, args, device: Optional[str] = None, rank: int = 0, world_size: int = 1):
    """``run_remove_background`` from the prepared dataset on (run.py:126-168, 228-330): inference, posterior, denoised
    counts per FPR.  File output, report and checkpoint tarballs stay with the caller.  Returns (posterior, {fpr: csc}).
    With ``world_size`` > 1 training is data parallel; every rank then holds the same model and computes the (small)
    remainder itself."""
    from .posterior import Posterior

    model, _, _, _ = run_inference(dataset_obj, args, device=device, rank=rank, world_size=world_size)
    model.eval()
    posterior = Posterior(dataset_obj=dataset_obj, vi_model=model, posterior_batch_size=args.posterior_batch_size,
                          debug=getattr(args, "debug", False))
    denoised = compute_output_denoised_counts(posterior, args)
    return posterior, denoised
