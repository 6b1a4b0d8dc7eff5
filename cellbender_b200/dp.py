"""Multi-GPU execution: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the single exchange
step of training; no collective on the posterior/estimation path until the final gather.  The reference is
single-GPU only ("CellBender only makes use of 1 GPU", docs/source/troubleshooting/index.rst:399), so this is new
surface; SURVEY.md section 8e states the semantics.

Training (data parallel): every rank holds the full parameter set and the surely-empty droplets, and a disjoint
shard of the analysed droplets (``shard_rows``).  Each step every rank runs its own 255-droplet minibatch (forward + backward as one CUDA-graph replay); the ELBO is
a plain SUM over droplets (train.py:56-63), so the exchange is ONE reduction of the flat fp32 gradient buffer
(69.4 M values at the PBMC8k shape), issued eagerly between graph replays: by default reduce-scatter -> ClippedAdam sweep
of this rank's 1/world_size shard -> all-gather of the parameters (``reduce_scatter_step``); ``CB_DP_MODE=allreduce``
selects all-reduce + the full sweep on every rank.
Stated consequences: the global minibatch is 255 x world_size droplets; an epoch is n_batches(shard) steps, so
OneCycleLR's total_steps shrinks by world_size; BatchNorm statistics, Dropout1d masks, epsilon medians and the
reparameterisation noise are per-rank (different CUDA seeds per rank; ``sync_buffers`` averages the BN running
statistics before evaluation); the encoder's first-forward offset heuristic
(encoder.py:300-309) is computed on rank 0 and broadcast.

Posterior / estimation: ``shard_cells`` splits the p > 0.5 cell list into contiguous ranges; ranks compute their COO
pieces independently (``Posterior.device_posterior(cell_subset=...)``) and ``gather_posterior`` concatenates them.
MCKP needs all cells of a gene, so it runs on the gathered COO (one gene-keyed exchange = the gather itself).
"""
import os
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world_size: int) -> np.ndarray:
    """Strided shard: droplets are sorted by UMI count, so striding keeps the cell / empty mix of every shard equal."""
    return np.arange(rank, n_rows, world_size)


def shard_cells(n_cells: int, rank: int, world_size: int) -> np.ndarray:
    """Contiguous range of the cell list (keeps each rank's COO piece in (m, c) order)."""
    bounds = np.linspace(0, n_cells, world_size + 1).astype(np.int64)
    return np.arange(bounds[rank], bounds[rank + 1])


def dp_mode() -> str:
    """Exchange mode of data-parallel training: 'p2p' (default: exchange fused into the optimizer sweep over NVLink peer
    memory, ``p2p_step``), 'sharded' (NCCL reduce_scatter -> sweep -> all_gather), 'allreduce', 'overlap'."""
    return os.environ.get("CB_DP_MODE", "p2p")


def p2p_group():
    """Process group for symmetric-memory (NVLink peer) parameter / gradient buffers, or None when the fused P2P
    exchange is not selected (CB_DP_MODE=p2p) or torch.distributed is not running on NCCL."""
    if dp_mode() != "p2p" or not dist.is_available() or not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    if dist.get_world_size() > 16:  # cb_clipped_adam_p2p carries at most 16 gradient sources
        return None
    return dist.group.WORLD


def attach(svi, world_size: int):
    """Make ``svi.step`` data-parallel: all-reduce(SUM) the flat gradient buffer between backward and the optimizer."""
    if world_size <= 1:
        return svi
    rank = dist.get_rank()
    torch.manual_seed(1234 + 104729 * rank)
    torch.cuda.manual_seed(1234 + 104729 * rank)
    svi.after_backward = lambda opt: dist.all_reduce(opt.flat_g, op=dist.ReduceOp.SUM)
    if getattr(getattr(svi.optim, "flat_g", None), "is_cuda", False):
        # measured on 2 x B200 (r01, ms/step): 1 piece 2.04, 2 pieces 2.10, 4 pieces 2.14 -- the Adam sweep is HBM-bound and
        # the NCCL kernels compete for the same HBM, so pipelining them buys nothing; one all-reduce is the default
        n_chunks = int(os.environ.get("CB_DP_CHUNKS", "1"))
        # measured on B200 (r01, ms/step at 2 / 4 GPUs): sharded 1.87 / 1.96, overlap 1.92 / 2.06, allreduce 1.87 / 2.10
        mode = dp_mode()
        if mode == "p2p" and (getattr(svi.optim, "symm_p", None) is None or svi.optim.n % (4 * world_size) != 0):
            # the optimizer's buffers are not NVLink peer memory (symmetric-memory rendezvous unavailable on this system,
            # optimizer built without a group, more than 16 ranks): NCCL carries the exchange instead
            if "CB_DP_MODE" in os.environ:
                raise RuntimeError("CB_DP_MODE=p2p needs the optimizer's flat buffers in symmetric memory "
                                   "(run.setup_inference / optim.get_optimizer(symmetric_group=...))")
            mode = "sharded"
        if mode == "p2p":
            # measured on B200 (r01, ms/step): 2 GPUs 1.46 against 1.85 for the NCCL exchange; 4 GPUs fused kernel 0.64 ms,
            # step ~1.72 (the NVLS variant with the same kernel time measured 1.714) against 1.96 (profiles/r01d_dp_exchange.txt)
            svi.exchange_and_step = lambda opt: p2p_step(opt, world_size, rank)
        elif mode == "overlap" and n_chunks <= 1:
            # optional: the reduction of each large gradient block is issued right behind its weight-gradient GEMM, on
            # that GEMM's side stream, followed by the ClippedAdam sweep of the block -- inside backward and inside the
            # captured step graph (NCCL collectives capture fine; drop the graphs before destroying the process group).
            # It hides the exchange behind backward but measures slower than the sharded sweep: the NCCL kernels and the
            # backward kernels compete for the same HBM / L2 bandwidth.
            svi.optim.dp_reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)
            svi.optim.early_weight_updates = True
            svi.dp_overlap = True
        elif n_chunks > 1:
            svi.exchange_and_step = lambda opt: allreduce_and_step(opt, n_chunks)
        elif mode == "sharded" and svi.optim.n % (4 * world_size) == 0:
            # default: reduce-scatter -> every rank sweeps 1/world_size of the flat buffers -> all-gather of the
            # parameters.  Same bytes on the NVLink fabric as the all-reduce, 1/world_size of the 1.9 GB optimizer sweep.
            svi.exchange_and_step = lambda opt: reduce_scatter_step(opt, world_size, rank)
    svi.sync_offsets = lambda: broadcast_encoder_offsets(svi.model)
    return svi


def chunk_bounds(n: int, n_chunks: int, align: int = 4):
    """[(begin, end)] covering [0, n) in n_chunks nearly equal pieces whose interior boundaries are multiples of
    ``align`` elements (16-byte alignment of the fp32 buffers for the vectorised Adam kernel)."""
    cuts = [0] + [min(n, (i * n // n_chunks) // align * align) for i in range(1, n_chunks)] + [n]
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


@torch.no_grad()
def allreduce_and_step(opt, n_chunks: int = 8):
    """The exchange step of data-parallel training, pipelined: the flat gradient buffer is all-reduced in ``n_chunks``
    pieces (queued back to back on NCCL's stream) and the ClippedAdam sweep of piece i runs on the compute stream as
    soon as piece i has arrived, i.e. while pieces i+1.. are still on the NVLink fabric.  Same arithmetic as
    ``all_reduce(flat_g)`` followed by ``opt.step()``."""
    from ._cabi import current_stream, lib

    if not opt.constant_learning_rate and opt.host_steps >= opt.total_steps:
        raise ValueError(f"Tried to step {opt.host_steps + 1} times. The specified number of total steps is {opt.total_steps}")
    bounds = chunk_bounds(opt.n, n_chunks)
    works = [dist.all_reduce(opt.flat_g[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in bounds]
    L, st = lib(), current_stream()
    tab = (opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), opt.lr_table.numel(), opt.step_count.data_ptr(),
           float(opt.beta2), float(opt.eps), float(opt.clip))
    for i, ((a, b), w) in enumerate(zip(bounds, works)):
        w.wait()  # makes the compute stream wait for this piece only
        at = lambda t: t.data_ptr() + 4 * a
        L.call("cb_clipped_adam_range", at(opt.flat_p), at(opt.flat_g), at(opt.flat_m), at(opt.flat_v), b - a, *tab,
               1 if i == len(bounds) - 1 else 0, st)
    opt.host_steps += 1


@torch.no_grad()
def reduce_scatter_step(opt, world_size: int, rank: int):
    """Data-parallel exchange with a sharded optimizer sweep: reduce_scatter(SUM) of the flat gradient buffer, ClippedAdam on
    this rank's contiguous shard of param / exp_avg / exp_avg_sq, all_gather of the updated parameter shards (in place).
    Arithmetic per element is that of ``all_reduce`` + ``opt.step()`` (NCCL's reduction order aside); only this rank's
    shard of the moment buffers is current -- ``gather_optimizer_state`` completes them before a checkpoint."""
    from ._cabi import current_stream, lib

    if not opt.constant_learning_rate and opt.host_steps >= opt.total_steps:
        raise ValueError(f"Tried to step {opt.host_steps + 1} times. The specified number of total steps is {opt.total_steps}")
    sh = opt.n // world_size
    a = rank * sh
    if getattr(opt, "_g_shard", None) is None or opt._g_shard.numel() != sh:
        opt._g_shard = torch.empty(sh, dtype=torch.float32, device=opt.flat_g.device)
    dist.reduce_scatter_tensor(opt._g_shard, opt.flat_g, op=dist.ReduceOp.SUM)
    at = lambda t: t.data_ptr() + 4 * a
    lib().call("cb_clipped_adam_range", at(opt.flat_p), opt._g_shard.data_ptr(), at(opt.flat_m), at(opt.flat_v), sh,
               opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), opt.lr_table.numel(), opt.step_count.data_ptr(),
               float(opt.beta2), float(opt.eps), float(opt.clip), 1, current_stream())
    dist.all_gather_into_tensor(opt.flat_p, opt.flat_p[a:a + sh])
    opt.host_steps += 1


def p2p_uses_nvls(opt, world_size: int) -> bool:
    """NVLS form (multimem.ld_reduce / multimem.st through the NVSwitch) when the symmetric buffers have multicast
    addresses and CB_P2P_NVLS=1 asks for it.  (Plain peer loads / stores put 2 (R - 1) / R parameter sets per step on
    every link, NVLS about one for any R.)"""
    have = bool(getattr(opt.symm_g, "has_multicast_support", False)) and int(getattr(opt.symm_g, "multicast_ptr", 0) or 0) != 0 \
        and int(getattr(opt.symm_p, "multicast_ptr", 0) or 0) != 0
    env = os.environ.get("CB_P2P_NVLS")
    if env == "0" or not have:
        if env == "1":
            raise RuntimeError("CB_P2P_NVLS=1 but the symmetric buffers have no multicast address on this system")
        return False
    # measured on B200 (r01): fused kernel 0.72 ms (NVLS) vs 0.43 ms (peer loads / stores) at 2 GPUs, 0.64 vs 0.64 at 4:
    # the one-load-per-thread NVLS loop is latency-bound, so it stays opt-in until it is unrolled
    return env == "1"


@torch.no_grad()
def p2p_step(opt, world_size: int, rank: int):
    """The data-parallel exchange fused into the optimizer sweep, over NVLink peer memory (cb_clipped_adam_p2p):
    device-side barrier (every rank has finished backward) -> ONE kernel that, for this rank's shard, sums the gradient
    shards of all ranks straight out of their gradient buffers, applies ClippedAdam and stores the new parameters into
    every rank's parameter buffer -> device-side barrier (every shard delivered).  Replaces NCCL reduce_scatter + sweep +
    all_gather; same arithmetic per element (gradients summed in rank order).  Moment buffers are sharded as in
    ``reduce_scatter_step``."""
    import ctypes

    from ._cabi import current_stream, lib

    if not opt.constant_learning_rate and opt.host_steps >= opt.total_steps:
        raise ValueError(f"Tried to step {opt.host_steps + 1} times. The specified number of total steps is {opt.total_steps}")
    sh = opt.n // world_size
    a = rank * sh
    cache = getattr(opt, "_p2p_ptrs", None)
    if cache is None:
        f32 = torch.float32
        grads = [opt.symm_g.get_buffer(r, (opt.n,), f32).data_ptr() + 4 * a for r in range(world_size)]
        peers = [opt.symm_p.get_buffer(r, (opt.n,), f32).data_ptr() + 4 * a for r in range(world_size) if r != rank]
        assert grads[rank] == opt.flat_g.data_ptr() + 4 * a, "symmetric gradient buffer is not where flat_g lives"
        cache = opt._p2p_ptrs = ((ctypes.c_void_p * world_size)(*grads), (ctypes.c_void_p * max(1, len(peers)))(*peers),
                                 len(peers))
    g_arr, p_arr, n_peer = cache
    timeout_ms = 20000  # a rank that never arrives traps the waiting kernels instead of hanging the GPUs
    timing = os.environ.get("CB_DP_TIMING") == "1"  # CUDA events around the three phases (diagnostics, see p2p_timing)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timing else None
    if timing:
        ev[0].record()
    opt.symm_g.barrier(channel=0, timeout_ms=timeout_ms)
    if timing:
        ev[1].record()
    at = lambda t: t.data_ptr() + 4 * a
    tab = (opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), opt.lr_table.numel(), opt.step_count.data_ptr(),
           float(opt.beta2), float(opt.eps), float(opt.clip), 1, current_stream())
    if p2p_uses_nvls(opt, world_size):
        # in-switch reduction of the gradients, multicast store of the parameters
        lib().call("cb_clipped_adam_nvls", at(opt.flat_p), at(opt.flat_m), at(opt.flat_v), sh,
                   int(opt.symm_g.multicast_ptr) + 4 * a, int(opt.symm_p.multicast_ptr) + 4 * a, *tab)
    else:
        lib().call("cb_clipped_adam_p2p", at(opt.flat_p), at(opt.flat_m), at(opt.flat_v), sh, ctypes.addressof(g_arr),
                   world_size, rank, ctypes.addressof(p_arr), n_peer, *tab)
    if timing:
        ev[2].record()
    opt.symm_p.barrier(channel=1, timeout_ms=timeout_ms)
    if timing:
        ev[3].record()
        opt.__dict__.setdefault("_p2p_events", []).append(ev)
    opt.host_steps += 1


def p2p_timing(opt, last: int = 50):
    """Mean milliseconds of (barrier before, fused kernel, barrier after) over the last recorded steps (CB_DP_TIMING=1)."""
    evs = getattr(opt, "_p2p_events", [])[-last:]
    if not evs:
        return None
    torch.cuda.synchronize()
    return tuple(float(np.mean([e[i].elapsed_time(e[i + 1]) for e in evs])) for i in range(3))


@torch.no_grad()
def gather_optimizer_state(opt, world_size: int, rank: int):
    """Make the moment buffers complete on every rank after sharded stepping (call before saving a checkpoint)."""
    if world_size <= 1:
        return
    sh = opt.n // world_size
    for buf in (opt.flat_m, opt.flat_v):
        dist.all_gather_into_tensor(buf, buf[rank * sh:(rank + 1) * sh].clone())


@torch.no_grad()
def sync_buffers(model, world_size: int):
    """Average the BatchNorm running statistics over ranks (they are updated from per-rank minibatches and are not part
    of the gradient exchange).  Call at the end of an epoch / before the posterior so every rank holds the same eval
    model; ``num_batches_tracked`` takes the maximum."""
    if world_size <= 1:
        return
    for _, buf in model.named_buffers():
        if buf.is_floating_point():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.div_(world_size)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.MAX)


def broadcast_encoder_offsets(model):
    enc = model.encoder["other"]
    vals = [0.0, 0.0] if enc.offset is None else [enc.offset["logit_p"], enc.offset["epsilon"]]
    t = torch.tensor(vals, dtype=torch.float64, device=model.device)
    dist.broadcast(t, src=0)
    enc.offset = {"logit_p": float(t[0]), "epsilon": float(t[1])}


def all_gather_varlen(t: torch.Tensor, world_size: int) -> torch.Tensor:
    """Concatenate 1-D tensors of different lengths from all ranks, in rank order (pads to the longest piece)."""
    if world_size <= 1:
        return t
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world_size)]
    dist.all_gather(sizes, n)
    n_max = int(max(int(s) for s in sizes))
    buf = torch.zeros(n_max, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    outs = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(outs, buf)
    return torch.cat([o[: int(s)] for o, s in zip(outs, sizes)])


def gather_posterior(piece, world_size: int):
    """All-gather the ranks' DevicePosteriorCOO pieces (variable length) and concatenate in rank order."""
    from .estimation import DevicePosteriorCOO

    if world_size <= 1:
        return piece
    g = lambda t: all_gather_varlen(t, world_size)
    out = DevicePosteriorCOO(g(piece.m), g(piece.c), g(piece.log_prob), g(piece.off_m), g(piece.off_v), piece.n_cols)
    out.ensure_sorted()
    if out.off_m.numel() > 1:
        o = torch.argsort(out.off_m)
        out.off_m, out.off_v = out.off_m[o].contiguous(), out.off_v[o].contiguous()
    return out
