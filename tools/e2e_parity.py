"""End-to-end parity of remove-background: product (B200 kernels) vs the reference-style pipeline.

north_star: "end-to-end cell probabilities and removed-count totals within 1% at a fixed seed".  The reference's own
integration test is tests/test_integration.py:15-65 (full CLI run on the simulated fixture of tests/conftest.py:77-105);
the pipeline it exercises is run.py:126-168 + 228-330.

    python tools/e2e_parity.py --make-golden --config c1 --epochs 150 --seeds 3     # build container (CPU, ~8 min / seed)
    python tools/e2e_parity.py --product --config c1 --epochs 150                  # GPU box: product vs tests/golden/e2e_c1.npz

--make-golden runs oracle/pipeline.py (training = the oracle restatement of SVI + TraceEnum_ELBO + ClippedAdam + OneCycleLR;
posterior log-prob / MCKP targets / MCKP through the UNMODIFIED reference functions imported via oracle/ref_shim.py when
/root/reference is present) for several seeds and stores, per seed: cell probabilities of the analysed droplets, per-gene and
total removed counts, per-cell removed counts, the learned scalars.  The seed-to-seed spread of the oracle itself is stored
too: it is the noise floor of any comparison between two stochastic trainings.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def summarize(ds, p, noise_csr, denoised_csc):
    """What is compared: cell calls, probabilities, removed counts (total / per gene / per called cell)."""
    cells = p > 0.5
    cell_inds = np.asarray(ds.analyzed_barcode_inds)[cells]
    raw_cells = ds.matrix[cell_inds]
    removed_gene = np.asarray(noise_csr.sum(axis=0)).ravel().astype(np.float64)
    removed_cell = np.asarray(noise_csr[np.asarray(ds.analyzed_barcode_inds)].sum(axis=1)).ravel().astype(np.float64)
    out = {"p": p.astype(np.float64), "n_cells": int(cells.sum()), "removed_total": float(removed_gene.sum()),
           "removed_per_gene": removed_gene, "removed_per_analyzed_droplet": removed_cell,
           "raw_total_in_cells": float(raw_cells.sum()), "denoised_total": float(denoised_csc.sum())}
    if "bkg_per_droplet" in ds.truth:
        out["true_background_in_cells"] = float(np.asarray(ds.truth["bkg_per_droplet"])[cell_inds].sum())
    elif "counts_bkg" in ds.truth:
        out["true_background_in_cells"] = float(ds.truth["counts_bkg"][cell_inds].sum())
    return out


def make_golden(a):
    import torch

    from cellbender_b200 import synth
    from oracle import pipeline as pl
    from oracle import ref_shim

    torch.set_num_threads(os.cpu_count())
    ds = synth.make_config_dataset(a.config, seed=0, host=True)
    # the dataset itself travels as a fixture: numpy's samplers are not bit-reproducible across CPUs (libm variants)
    cfg = synth.CONFIGS[a.config]
    synth.save_fixture_dataset(os.path.join(GOLDEN_DIR, f"e2e_{a.config}_dataset.npz"), ds, cfg["expected_cells"], cfg["total_droplets"])
    ds = synth.load_fixture_dataset(os.path.join(GOLDEN_DIR, f"e2e_{a.config}_dataset.npz"))
    ref = None
    log_prob_fn = None
    if ref_shim.reference_available() and not a.no_reference:
        ref_shim.install()
        import cellbender.remove_background.estimation as ref_est
        import cellbender.remove_background.posterior as ref_post

        ref = (ref_est, ref_post)
        log_prob_fn = ref_post.Posterior._log_prob_noise_count_tensor
    runs = []
    for k in range(a.seeds):
        seed = 1234 + 1000 * k
        t0 = time.time()
        st = pl.train(ds, epochs=a.epochs, seed=seed, log=print)
        lat = pl.latents_map(ds, st)
        coo, offsets = pl.posterior_coo(ds, st, lat["p"], seed=seed, log_prob_fn=log_prob_fn)
        den = pl.denoise(ds, coo, offsets, lat["p"], fpr=0.01, reference_estimation=ref)
        s = summarize(ds, lat["p"], den["noise_csr"], den["denoised_csc"])
        _, ps = pl._split(st["params"], st["buffers"])
        s["scalars"] = {k_: float(v.reshape(-1)[0]) for k_, v in ps.items() if v.numel() == 1}
        s["chi_ambient"] = ps["chi_ambient"].numpy().astype(np.float64)
        s["train_loss_last"] = st["train_loss"][-1]
        s["coo_nnz"] = int(coo.nnz)
        s["seconds"] = time.time() - t0
        print(f"seed {seed}: cells {s['n_cells']}, removed {s['removed_total']:.0f} of {s['raw_total_in_cells']:.0f} "
              f"(true background {s.get('true_background_in_cells', float('nan')):.0f}), loss {s['train_loss_last']:.3f}, "
              f"{s['seconds']:.0f} s", flush=True)
        runs.append(s)
    out = {"config": a.config, "epochs": a.epochs, "dataset_checksum": np.uint64(synth.dataset_checksum(ds.matrix)),
           "seeds": np.array([1234 + 1000 * k for k in range(a.seeds)]),
           "reference_functions_used": np.array(ref is not None),
           "p": np.stack([r["p"] for r in runs]), "removed_per_gene": np.stack([r["removed_per_gene"] for r in runs]),
           "removed_per_analyzed_droplet": np.stack([r["removed_per_analyzed_droplet"] for r in runs]),
           "removed_total": np.array([r["removed_total"] for r in runs]), "n_cells": np.array([r["n_cells"] for r in runs]),
           "raw_total_in_cells": np.array([r["raw_total_in_cells"] for r in runs]),
           "true_background_in_cells": np.array([r.get("true_background_in_cells", np.nan) for r in runs]),
           "chi_ambient": np.stack([r["chi_ambient"] for r in runs]),
           "train_loss_last": np.array([r["train_loss_last"] for r in runs]),
           "scalars_json": json.dumps([r["scalars"] for r in runs])}
    path = os.path.join(GOLDEN_DIR, f"e2e_{a.config}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def run_product(config: str, epochs: int, device: str = "cuda:0", fpr: float = 0.01, log=None, rank: int = 0,
                world_size: int = 1):
    """The product end to end through its public API: run_inference -> Posterior.latents_map ->
    cell_noise_count_posterior_coo -> compute_mean_target_removal_as_function -> compute_denoised_counts(MCKP) ->
    restore_eliminated_features_in_cells (cellbender_b200.run.run_remove_background)."""
    import torch

    from cellbender_b200 import run, synth

    ds = synth.load_fixture_dataset(os.path.join(GOLDEN_DIR, f"e2e_{config}_dataset.npz"))
    args = run.default_args(epochs=epochs, fpr=[fpr])
    t0 = time.time()
    posterior, denoised = run.run_remove_background(ds, args, device=device, rank=rank, world_size=world_size)
    torch.cuda.synchronize()
    seconds = time.time() - t0
    p = posterior.latents_map["p"]
    den = denoised[fpr]
    cells = p > 0.5
    cell_inds = np.asarray(ds.analyzed_barcode_inds)[cells]
    keep = np.zeros(ds.matrix.shape[0], dtype=bool)
    keep[cell_inds] = True
    import scipy.sparse as sp

    noise = (sp.csr_matrix(sp.diags(keep.astype(ds.matrix.dtype)) @ ds.matrix) - den.tocsr()).tocsr()
    s = summarize(ds, p, noise, den)
    m = posterior.vi_model
    s["scalars"] = {k: float(m.param_store[k].detach().reshape(-1)[0]) for k in m.param_store.keys()
                    if m.param_store[k].numel() == 1}
    s["chi_ambient"] = m.param_store["chi_ambient"].detach().double().cpu().numpy()
    s["train_loss_last"] = -m.loss["train"]["elbo"][-1]
    s["seconds"] = seconds
    s["coo_nnz"] = int(posterior._noise_count_posterior_coo.nnz)
    return ds, s


def compare(golden, s, checksum: int):
    """Product summary vs the golden runs.  Returns a dict of relative differences (vs the mean over golden seeds) and the
    golden seed-to-seed spread of each quantity."""
    assert int(golden["dataset_checksum"]) == int(checksum), "dataset differs from the one the golden run saw"
    gp = golden["p"]
    p_ref = gp.mean(0)
    calls_ref = (gp > 0.5)
    stable = calls_ref.all(0) | (~calls_ref).all(0)  # droplets every golden seed calls the same way
    rel = lambda x, y: float(abs(x - y) / max(abs(y), 1e-12))
    rt = golden["removed_total"]
    gene_ref = golden["removed_per_gene"].mean(0)
    big = gene_ref >= 0.01 * gene_ref.sum()  # genes carrying at least 1 % of the removed counts
    out = {
        "n_cells": int(s["n_cells"]), "n_cells_golden": [int(v) for v in golden["n_cells"]],
        "cell_calls_equal_on_stable_droplets": bool(((s["p"] > 0.5) == calls_ref[0])[stable].all()),
        "n_unstable_droplets_golden": int((~stable).sum()),
        "p_max_abs_diff_on_stable": float(np.abs(s["p"] - p_ref)[stable].max()),
        "p_golden_seed_spread_max": float((gp.max(0) - gp.min(0))[stable].max()),
        "removed_total": s["removed_total"], "removed_total_golden": [float(v) for v in rt],
        "removed_total_rel_diff": rel(s["removed_total"], rt.mean()),
        "removed_total_golden_seed_spread_rel": float((rt.max() - rt.min()) / rt.mean()),
        "removed_per_gene_rel_diff_l1": float(np.abs(s["removed_per_gene"] - gene_ref).sum() / gene_ref.sum()),
        "removed_per_gene_golden_seed_spread_l1": float(
            max(np.abs(g - gene_ref).sum() for g in golden["removed_per_gene"]) / gene_ref.sum()),
        "removed_big_genes_max_rel_diff": float((np.abs(s["removed_per_gene"] - gene_ref)[big] / gene_ref[big]).max()) if big.any() else 0.0,
        "n_big_genes": int(big.sum()),
        "chi_ambient_l1_diff": float(np.abs(s["chi_ambient"] - golden["chi_ambient"].mean(0)).sum()),
        "chi_ambient_golden_seed_spread_l1": float(max(np.abs(c - golden["chi_ambient"].mean(0)).sum() for c in golden["chi_ambient"])),
        "train_loss_last": s["train_loss_last"], "train_loss_last_golden": [float(v) for v in golden["train_loss_last"]],
        "scalars": s["scalars"], "scalars_golden": json.loads(str(golden["scalars_json"])),
        "true_background_in_cells": s.get("true_background_in_cells"),
    }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--make-golden", action="store_true")
    ap.add_argument("--product", action="store_true")
    ap.add_argument("--config", default="c1")
    ap.add_argument("--epochs", type=int, default=150)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.make_golden:
        make_golden(a)
    if a.product:
        from cellbender_b200 import synth

        golden = np.load(os.path.join(GOLDEN_DIR, f"e2e_{a.config}.npz"), allow_pickle=False)
        assert int(golden["epochs"]) == a.epochs, "golden was made with a different number of epochs"
        # under torchrun: data-parallel training over all ranks (global minibatch 255 x world_size, total steps / world_size);
        # every rank finishes the pipeline with the common model, rank 0 reports
        world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
        if world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        ds, s = run_product(a.config, a.epochs, device=f"cuda:{local}", rank=rank, world_size=world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
            if rank != 0:
                return
        res = compare(golden, s, synth.dataset_checksum(ds.matrix))
        res["product_seconds"] = s["seconds"]
        res["world_size"] = world
        text = json.dumps(res, indent=1)
        print(text)
        if a.out:
            open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
