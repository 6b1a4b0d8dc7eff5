"""N > 1 host logic on the CPU with gloo, world_size 2: shard arithmetic, the gradient exchange (SUM all-reduce of the
flat buffer), the encoder-offset broadcast and the variable-length gather of posterior pieces.  (The kernels themselves
need a GPU; the multi-GPU numbers come from `gpurun --gpus N` runs of bench.py / tests/test_multi_gpu.py.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cellbender_b200 import dp


def test_shards_are_disjoint_and_complete():
    for n, w in ((22500, 2), (22500, 8), (7, 4), (8000, 3)):
        rows = [dp.shard_rows(n, r, w) for r in range(w)]
        cells = [dp.shard_cells(n, r, w) for r in range(w)]
        for parts in (rows, cells):
            cat = np.concatenate(parts)
            assert cat.size == n and np.array_equal(np.sort(cat), np.arange(n))
        assert all(np.all(np.diff(c) == 1) for c in cells if c.size > 1)  # contiguous -> COO pieces stay (m, c)-sorted
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length gather keeps rank order
        t = torch.arange(3 + 2 * rank, dtype=torch.int64) + 100 * rank
        got = dp.all_gather_varlen(t, world)
        # gradient exchange: SUM of the flat buffers
        class Opt:  # the two attributes the hook touches
            flat_g = torch.full((5,), float(rank + 1))

        class Svi:
            optim = Opt()
            after_backward = None

        svi = dp.attach(Svi(), world)
        svi.after_backward(svi.optim)
        # encoder offsets broadcast from rank 0
        class Enc:
            offset = {"logit_p": 1.5, "epsilon": -0.25} if rank == 0 else None

        class Model:
            device = "cpu"
            encoder = {"other": Enc()}

        m = Model()
        dp.broadcast_encoder_offsets(m)
        # ranks agree on the number of steps per epoch (the shard loaders' lengths differ by one) and on the test ELBO
        steps = dp.agree_on_steps(33 + rank)
        mean = dp.all_reduce_mean(10.0 + 4.0 * rank, "cpu")
        q.put((rank, got.tolist(), Svi.optim.flat_g.tolist(), m.encoder["other"].offset, steps, mean))
    finally:
        dist.destroy_process_group()


def test_region_shards_tile_every_range_with_aligned_cuts():
    """dp.region_shard: the R shards of a range of the flat buffers tile it exactly, in rank order, on 16-byte boundaries,
    for lengths that are not multiples of 4 R (the tail behind the large blocks) and ranges shorter than R float4s."""
    for off, n in ((0, 69_547_264), (17_187_840, 4 * 1013), (64, 4), (128, 12), (0, 4096)):
        for world in (1, 2, 3, 4, 8, 16):
            cuts = [dp.region_shard(off, n, r, world) for r in range(world)]
            assert cuts[0][0] == off and cuts[-1][1] == off + n
            assert all(x[1] == y[0] for x, y in zip(cuts[:-1], cuts[1:]))
            assert all(a % 4 == 0 and b % 4 == 0 and b >= a for a, b in cuts)
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 4


def test_two_rank_exchange_on_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    expect = list(range(3)) + [100 + i for i in range(5)]
    for rank, gathered, flat_g, offset, steps, mean in res:
        assert gathered == expect
        assert flat_g == [3.0] * 5  # 1 + 2
        assert offset == {"logit_p": 1.5, "epsilon": -0.25}
        assert steps == 33 and mean == 12.0


def test_owner_shard_exchange_scheme_equals_allreduce_then_full_sweep():
    """The arithmetic contract of dp.p2p_step / dp.reduce_scatter_step, emulated in numpy with the oracle's ClippedAdam:
    every rank owns one contiguous shard, sums the shard's gradients of all ranks in rank order, updates its shard and
    delivers the new parameters to every rank.  Result: every replica holds exactly what all-reduce + a full sweep on
    every rank would give (the sum order per element is the same rank order), for several steps and rank counts."""
    from oracle.train_step import ClippedAdamPerTensor

    rng = np.random.default_rng(0)
    n = 64 * 12
    for world in (2, 3, 4, 8):
        if n % (4 * world):
            continue
        p0 = torch.tensor(rng.normal(size=n), dtype=torch.float32)
        # reference: one replica, gradients summed in rank order, full sweep
        ref_p = {"w": p0.clone()}
        ref_opt = ClippedAdamPerTensor(ref_p, lr=1e-3)
        # scheme: world replicas of the parameters, but moments only for the owned shard
        sh = n // world
        replicas = [p0.clone() for _ in range(world)]
        shard_p = [{"w": replicas[r][r * sh:(r + 1) * sh].clone()} for r in range(world)]
        shard_opt = [ClippedAdamPerTensor(shard_p[r], lr=1e-3) for r in range(world)]
        for step in range(3):
            grads = [torch.tensor(rng.normal(scale=8.0, size=n), dtype=torch.float32) for _ in range(world)]
            total = grads[0].clone()
            for g in grads[1:]:
                total += g
            ref_p["w"].grad = total.clone()
            ref_opt.step()
            for r in range(world):  # owner r: rank-order sum of its shard, sweep, deliver to everybody
                acc = grads[0][r * sh:(r + 1) * sh].clone()
                for g in grads[1:]:
                    acc += g[r * sh:(r + 1) * sh]
                shard_p[r]["w"].grad = acc
                shard_opt[r].step()
                for rep in replicas:
                    rep[r * sh:(r + 1) * sh] = shard_p[r]["w"]
            for rep in replicas:
                assert torch.equal(rep, replicas[0])
            assert torch.equal(replicas[0], ref_p["w"]), (world, step)
