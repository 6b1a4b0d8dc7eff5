"""-m gpu, ONE device: the data-parallel host path end to end with R virtual ranks on one GPU (the driver's box has a single
GPU, where tests/test_multi_gpu.py is skipped).  Every virtual rank is a complete training set-up (model, flat optimizer,
SVI, shard loader) built by run.setup_inference(rank=r, world_size=R); their "peer" buffers are each other's local tensors
(dp.LocalPeers).  Each step runs forward + backward on every virtual rank and then dp.p2p_step for every rank -- the same
Python glue and the same kernel (cb_clipped_adam_p2p) the multi-GPU run uses -- and is checked against the arithmetic
contract of the exchange: all-reduce(SUM) of the gradients followed by the full ClippedAdam sweep."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dataset(n_included=501):
    from cellbender_b200 import synth

    sim = synth.simulate_dataset(n_genes=256, cells_of_each_type=[100, 100], n_droplets=1500, cell_mean_umi=2000.0,
                                 empty_mean_umi=60.0, dirichlet_alpha=0.5, seed=0, device="cpu", cell_chunk=128)
    return synth.prepare_dataset(sim, expected_cells=200, total_droplets_included=n_included)


@pytest.mark.parametrize("world", [2, 4])
def test_virtual_ranks_p2p_step_equals_allreduce_plus_full_sweep(world):
    from cellbender_b200 import dp, ops, run

    ds = _dataset()
    args = run.default_args(epochs=4, z_dim=8, z_hidden_dims=[32], learning_rate=1e-3, use_cuda_graph=False)
    ranks = [run.setup_inference(ds, args, device=DEV, rank=r, world_size=world) for r in range(world)]
    models, opts, svis, loaders = [x[0] for x in ranks], [x[1] for x in ranks], [x[2] for x in ranks], [x[3] for x in ranks]
    # uneven shards (501 droplets over 2 or 4 ranks, random train/test split per rank): agree on the epoch length by hand
    # (dp.agree_on_steps needs a process group) -- the optimizers must have been built with the same total_steps
    n_steps = min(len(tl) for tl in loaders)
    for tl in loaders:
        tl.limit_batches(n_steps)
        assert len(tl) == n_steps and sum(1 for _ in tl) == n_steps
    assert len({o.n for o in opts}) == 1 and opts[0].n % (4 * world) == 0
    for o in opts:
        assert torch.equal(o.flat_p, opts[0].flat_p)  # same seed -> identical initial replicas
        o.symm_g = dp.LocalPeers([q.flat_g for q in opts])
        o.symm_p = dp.LocalPeers([q.flat_p for q in opts])
        o.set_total_steps(n_steps * args.epochs)  # what dp.agree_on_steps gives every rank in a real run
    for r, (svi, m) in enumerate(zip(svis, models)):
        svi.after_backward = lambda o: None  # data-parallel marker: no early block updates, the exchange is explicit below
        m.train()
    # reference state of the contract: full-size moment buffers swept with the summed gradient
    p_ref, m_ref, v_ref = opts[0].flat_p.clone(), torch.zeros_like(opts[0].flat_p), torch.zeros_like(opts[0].flat_p)
    step_ref = torch.zeros(1, dtype=torch.int64, device=DEV)
    # three ranges with sizes that are multiples of 4 floats but not of 4 x world: shards of every range are uneven
    n = opts[0].n
    cuts = [0, 4 * 1013, 4 * 1013 + 4 * 2027, n]
    regions = [(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])] if n > cuts[2] + 64 else None
    for epoch in range(2):
        its = [iter(tl) for tl in loaders]
        for k in range(n_steps):
            for r in range(world):
                svis[r]._forward_backward(next(its[r]))
            g_sum = opts[0].flat_g.clone()
            for o in opts[1:]:
                g_sum += o.flat_g
            ops.clipped_adam_step(p_ref, g_sum, m_ref, v_ref, opts[0].lr_table, opts[0].beta1_table, step_ref)
            for r in range(world):
                dp.p2p_step(opts[r], world, r, regions=regions)
            torch.cuda.synchronize()
            covered = torch.zeros(n, dtype=torch.int32, device=DEV)
            for r, o in enumerate(opts):
                assert torch.equal(o.flat_p, opts[0].flat_p), "replicas diverged"
                for a, b in o.exchange.owned:  # the moment buffers are current where the rank owns the shard
                    torch.testing.assert_close(o.flat_m[a:b], m_ref[a:b], rtol=1e-6, atol=1e-9)
                    torch.testing.assert_close(o.flat_v[a:b], v_ref[a:b], rtol=1e-6, atol=1e-12)
                    covered[a:b] += 1
                assert int(o.step_count.item()) == int(step_ref.item()) == o.host_steps
            assert bool((covered == 1).all()), "the ranks' shards must tile the flat buffers exactly once"
            torch.testing.assert_close(opts[0].flat_p, p_ref, rtol=1e-6, atol=1e-7)
        for it in its:  # every rank's epoch ends after the agreed number of steps
            with pytest.raises(StopIteration):
                next(it)


def test_loader_serves_the_agreed_number_of_batches():
    """dataprep.DataLoader.limit_batches: the epoch ends after n minibatches and the next epoch starts reshuffled; without a
    limit the reference's behaviour (all droplets, trailing < 4 dropped) is untouched."""
    from cellbender_b200 import run

    ds = _dataset(500)
    args = run.default_args(epochs=1, z_dim=8, z_hidden_dims=[32])
    _, _, _, tl, _ = run.setup_inference(ds, args, device=DEV)
    full = len(tl)
    assert full >= 3
    tl.limit_batches(full - 1)
    assert len(tl) == full - 1
    a = [b.row_ids_host.copy() for b in tl]
    b = [b.row_ids_host.copy() for b in tl]
    assert len(a) == len(b) == full - 1
    assert not all(np.array_equal(x, y) for x, y in zip(a, b))  # reshuffled between epochs
