"""-m gpu, needs >= 2 GPUs (skipped otherwise): data-parallel training keeps replicas identical; barcode-sharded
posterior + gather equals the single-GPU posterior."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _dataset():
    from cellbender_b200 import synth

    sim = synth.simulate_dataset(n_genes=256, cells_of_each_type=[100, 100], n_droplets=1500, cell_mean_umi=2000.0,
                                 empty_mean_umi=60.0, dirichlet_alpha=0.5, seed=0, device="cpu", cell_chunk=128)
    return synth.prepare_dataset(sim, expected_cells=200, total_droplets_included=501)


def _worker(rank, world, port, q, mode="sharded", posterior=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        from cellbender_b200 import dp, run, train
        from cellbender_b200.posterior import Posterior

        ds = _dataset()  # 501 analysed droplets: uneven shards, the shard loaders' lengths differ between ranks
        args = run.default_args(epochs=3, z_dim=8, z_hidden_dims=[32], learning_rate=1e-3)
        model, opt, svi, tl, te = run.setup_inference(ds, args, device=dev, rank=rank, world_size=world, dp_mode=mode)
        dp.attach(svi, world, mode=mode)
        assert svi.dp_mode == mode
        p0 = opt.flat_p.clone()
        # the reference's training loop, unchanged: offsets broadcast after the first step, BatchNorm running statistics
        # averaged at the end of every epoch, the test ELBO averaged over ranks, every rank takes the agreed number of steps
        train.run_training(model, svi, tl, te, epochs=3, test_freq=1)
        torch.cuda.synchronize()
        assert opt.host_steps == 3 * len(tl) == int(opt.step_count.item())
        offs = torch.tensor([model.encoder["other"].offset["logit_p"]], device=dev, dtype=torch.float64)
        all_offs = [torch.zeros_like(offs) for _ in range(world)]
        dist.all_gather(all_offs, offs)
        assert all(torch.equal(all_offs[0], o) for o in all_offs), "encoder offsets differ between ranks"
        bn = model.encoder["other"].batchnorm0.running_mean.detach().clone()
        all_bn = [torch.zeros_like(bn) for _ in range(world)]
        dist.all_gather(all_bn, bn)
        assert all(torch.equal(all_bn[0], o) for o in all_bn), "BatchNorm running statistics differ between ranks"
        sd = opt.state_dict()  # gathers the sharded moments: identical on every rank afterwards
        fm = sd["flat_m"].detach().clone()
        all_m = [torch.zeros_like(fm) for _ in range(world)]
        dist.all_gather(all_m, fm)
        assert all(torch.equal(all_m[0], o) for o in all_m) and float(fm.abs().max()) > 0
        flat = opt.flat_p.detach().clone()
        others = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(others, flat)
        same = all(torch.equal(others[0], o) for o in others)
        moved = float((flat - p0).abs().max())
        if not posterior:
            q.put((rank, same, moved, flat.cpu().numpy(), len(svi._graphs)))
            svi._graphs.clear()
            torch.cuda.synchronize()
            return
        # posterior: shards + gather vs everything on one rank, seeded per chunk, shard boundaries on chunk boundaries
        model.eval()
        post = Posterior(ds, model, cells_per_chunk=32)
        n_cells = 128
        post._latents = {"p": np.concatenate([np.ones(n_cells), np.zeros(ds.analyzed_barcode_inds.size - n_cells)])}
        piece = post.device_posterior(cell_subset=dp.shard_cells(n_cells, rank, world, align=32), seed=7, n_samples=4)
        full = dp.gather_posterior(piece, world)
        # MCKP with the gene-keyed exchange (all-to-all of the pieces by gene owner) vs MCKP on the gathered COO
        from cellbender_b200.estimation import mckp_device

        n_genes_total = post.index_converter.total_n_genes
        targets = np.random.default_rng(11).uniform(0.0, 40.0, size=n_genes_total)
        seg_m, noise = dp.mckp_gene_sharded(piece, targets, n_genes_total, world, rank)
        ref_m, ref_noise = mckp_device(full, targets, n_genes_total)
        ok_mckp = torch.equal(seg_m, ref_m) and torch.equal(noise, ref_noise) and int(noise.sum()) > 0
        ok_post = None
        if rank == 0:
            ref = post.device_posterior(seed=7, n_samples=4)
            ok_post = (torch.equal(ref.m, full.m) and torch.equal(ref.c, full.c) and torch.equal(ref.log_prob, full.log_prob)
                       and torch.equal(ref.off_m, full.off_m) and torch.equal(ref.off_v, full.off_v) and ref.m.numel() > 1000)
        ok_post = ok_mckp if ok_post is None else (ok_post and ok_mckp)
        q.put((rank, same, moved, ok_post, len(svi._graphs)))
        svi._graphs.clear()  # graphs that captured NCCL collectives must be gone before the process group is destroyed
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_replicas_stay_identical_and_sharded_posterior_matches():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=150) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, same, moved, ok_post, n_graphs in res:
        assert same, "parameter replicas diverged"
        assert moved > 0
        assert n_graphs >= 1  # forward + backward ran as CUDA-graph replays around the eager all-reduce
    assert all(r[3] is True for r in res)  # sharded posterior == unsharded (rank 0), gene-sharded MCKP == gathered MCKP


def _train(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + (7 if mode == "p2p" else 3)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, False)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=150) for _ in range(world)), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs), mode
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_p2p_exchange_matches_nccl_sharded_exchange():
    """mode 'p2p' (cb_clipped_adam_p2p: gradient shards summed out of the peers' buffers over NVLink, ClippedAdam,
    parameters stored into every rank's buffer, all in one kernel between two device-side barriers) must train exactly
    like the NCCL reduce_scatter -> sweep -> all_gather exchange: replicas identical, same parameters as the NCCL run."""
    ref = _train(2, "sharded")
    got = _train(2, "p2p")
    for rank, same, moved, flat, n_graphs in got:
        assert same, "parameter replicas diverged with the fused P2P exchange"
        assert moved > 0 and n_graphs >= 1
    np.testing.assert_allclose(got[0][3], ref[0][3], rtol=0, atol=1e-6)
