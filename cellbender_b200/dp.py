"""Multi-GPU execution: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing; the single
exchange step of training is ONE kernel over NVLink peer memory; no collective on the posterior/estimation path until
the final gather.  The reference is single-GPU only ("CellBender only makes use of 1 GPU",
docs/source/troubleshooting/index.rst:399), so this is new surface; SURVEY.md section 8e states the semantics.

Training (data parallel): every rank holds the full parameter set and the surely-empty droplets, and a disjoint
shard of the analysed droplets (``shard_rows``).  Each step every rank runs its own 255-droplet minibatch (forward +
backward as one CUDA-graph replay); the ELBO is a plain SUM over droplets (train.py:56-63), so the exchange is ONE
reduction of the flat fp32 gradient buffer (69.4 M values at the PBMC8k shape) followed by the optimizer sweep.  Two
exchange paths, chosen once in ``attach`` (no environment switch):
  'p2p'      (default) the exchange is FUSED into the optimizer sweep over NVLink peer memory (``P2PExchange``,
             cb_clipped_adam_p2p / cb_clipped_adam_nvls): of every range of the flat buffers each rank owns a 1/R shard,
             sums the peers' gradient shards out of their buffers (TMA bulk reads, or one multimem.ld_reduce through the
             NVSwitch), applies ClippedAdam and stores the new parameters into every rank's buffer.  The large weight
             blocks are exchanged from inside backward, right behind their weight-gradient products, so the NVLink
             transfer overlaps the rest of backward; the whole step, barriers included, is ONE captured graph;
  'sharded'  NCCL reduce_scatter -> sweep of the 1/R shard -> all_gather: the path taken when the symmetric-memory
             rendezvous is unavailable on the system.
Stated consequences: the global minibatch is 255 x world_size droplets; every rank takes the SAME number of steps per
epoch (``agree_on_steps``: the minimum over ranks of the shard loaders' lengths), so OneCycleLR's total_steps =
agreed steps x epochs shrinks by about world_size; BatchNorm batch statistics, Dropout1d masks, epsilon medians and the
reparameterisation noise are per rank (different CUDA seeds per rank); BatchNorm RUNNING statistics are averaged over
ranks at the end of every epoch and before evaluation (``sync_buffers``); the encoder's first-forward offset heuristic
(encoder.py:300-309) is computed on rank 0 and broadcast right after the first step; the test ELBO used by the failure
criteria is averaged over ranks; ``FlatClippedAdam.state_dict()`` gathers the sharded moment buffers.

Posterior / estimation: ``shard_cells`` splits the barcode-sorted cell list into contiguous ranges; ranks compute their
COO pieces independently (``Posterior.device_posterior(cell_subset=...)``) and ``gather_posterior`` concatenates them in
rank order, which IS the (m, c)-sorted whole.  MCKP needs all cells of a gene: ``mckp_gene_sharded`` re-keys the pieces by
gene with ONE all-to-all (rank r receives every entry of the genes g with g mod R == r, still (m, c)-sorted because the
pieces arrive in barcode order), solves its genes' knapsacks and the per-(cell, gene) results are gathered.
"""
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world_size: int) -> np.ndarray:
    """Strided shard: droplets are sorted by UMI count, so striding keeps the cell / empty mix of every shard equal."""
    return np.arange(rank, n_rows, world_size)


def shard_cells(n_cells: int, rank: int, world_size: int, align: int = 1) -> np.ndarray:
    """Contiguous range of the (barcode-sorted) cell list, boundaries on multiples of ``align`` (= cells_per_chunk makes
    sharded and unsharded posterior passes use the same chunks, hence the same per-chunk random streams)."""
    bounds = np.linspace(0, n_cells, world_size + 1).astype(np.int64)
    if align > 1:
        bounds[1:-1] = np.minimum(n_cells, (bounds[1:-1] + align - 1) // align * align)
    return np.arange(bounds[rank], bounds[rank + 1])


def p2p_group():
    """Process group for symmetric-memory (NVLink peer) parameter / gradient buffers, or None when torch.distributed is
    not running on NCCL (or spans more ranks than cb_clipped_adam_p2p carries gradient sources for)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    if dist.get_world_size() > 16:
        return None
    return dist.group.WORLD


def agree_on_steps(n_local: int, device=None) -> int:
    """Steps per epoch every rank will take: the minimum over ranks of the shard loaders' lengths.  The train / test split
    and the shard sizes differ by rank, so the loaders' lengths can differ by one; ranks that issued different numbers of
    exchange steps would dead-lock the barriers (and sweep their parameter shards with different OneCycleLR tables)."""
    if not dist.is_initialized() or dist.get_world_size() <= 1:
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def attach(svi, world_size: int, mode: Optional[str] = None):
    """Make ``svi.step`` data-parallel.  ``mode``: None = 'p2p' when the optimizer's flat buffers are NVLink peer memory
    (run.setup_inference arranges that when the symmetric-memory rendezvous works), else 'sharded'."""
    if world_size <= 1:
        return svi
    rank = dist.get_rank()
    torch.manual_seed(1234 + 104729 * rank)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(1234 + 104729 * rank)
    opt = svi.optim
    have_symm = getattr(opt, "symm_p", None) is not None and getattr(opt, "n", 1) % (4 * world_size) == 0
    if mode is None:
        mode = "p2p" if have_symm else "sharded"
    if mode == "p2p":
        if not have_symm:
            raise RuntimeError("data-parallel mode 'p2p' needs the optimizer's flat buffers in symmetric memory "
                               "(run.setup_inference / optim.get_optimizer(symmetric_group=...))")
        # the exchange runs INSIDE the step (and its captured graph): large blocks behind their weight-gradient products,
        # the rest in FlatClippedAdam.step -- the training loop sees an ordinary single-process step
        opt.exchange = P2PExchange(opt, world_size, rank)
    elif mode == "sharded":
        if getattr(opt, "n", 0) % (4 * world_size) == 0 and getattr(getattr(opt, "flat_g", None), "is_cuda", False):
            svi.exchange_and_step = lambda o: reduce_scatter_step(o, world_size, rank)
        else:  # host-logic tests (gloo): plain SUM all-reduce of the flat gradient buffer, then the full sweep
            svi.after_backward = lambda o: dist.all_reduce(o.flat_g, op=dist.ReduceOp.SUM)
    else:
        raise ValueError(f"unknown data-parallel mode {mode!r}: 'p2p' or 'sharded'")
    if mode == "sharded" and svi.after_backward is None:
        svi.after_backward = lambda o: None  # marks the step as split: the NCCL exchange runs outside the step graph
    svi.dp_mode = mode
    # consistency hooks the training loop calls (train.SVI.step / train.run_training / optim.state_dict)
    svi.sync_offsets = lambda: broadcast_encoder_offsets(svi.model)
    svi.sync_buffers = lambda: sync_buffers(svi.model, world_size)
    svi.reduce_mean = lambda v: all_reduce_mean(v, svi.model.device)
    if hasattr(opt, "flat_m"):
        opt.gather_state = lambda: gather_optimizer_state(opt, world_size, rank)
    return svi


class LocalPeers:
    """Single-device stand-in for the symmetric-memory handle of R ranks: rank r's "peer" buffers are the other virtual
    ranks' local tensors and the barriers are no-ops, because the virtual ranks are stepped one after the other on one
    stream (each owner only writes its own shard, so sequential execution equals the concurrent one).  Lets the whole
    ``p2p_step`` path -- pointer tables, shard arithmetic, cb_clipped_adam_p2p -- run on a one-GPU box
    (tests/test_dp_single_gpu.py)."""

    has_multicast_support = False
    multicast_ptr = 0

    def __init__(self, buffers: List[torch.Tensor]):
        self.buffers = buffers

    def get_buffer(self, rank: int, shape, dtype):
        return self.buffers[rank]

    def barrier(self, channel: int = 0, timeout_ms: int = 0):
        return None


def all_reduce_mean(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item()) / dist.get_world_size()


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
    st_tab = getattr(opt, "step_table", None)
    lib().call("cb_clipped_adam_range", at(opt.flat_p), opt._g_shard.data_ptr(), at(opt.flat_m), at(opt.flat_v), sh,
               opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), None if st_tab is None else st_tab.data_ptr(),
               opt.lr_table.numel(), opt.step_count.data_ptr(),
               float(opt.beta2), float(opt.eps), float(opt.clip), 1, current_stream())
    dist.all_gather_into_tensor(opt.flat_p, opt.flat_p[a:a + sh])
    opt.host_steps += 1


def p2p_uses_nvls(opt, world_size: int) -> bool:
    """NVLS form (multimem.ld_reduce / multimem.st through the NVSwitch) when the symmetric buffers have multicast
    addresses and there are at least ``NVLS_MIN_RANKS`` ranks.  Plain peer loads / stores put 2 (R - 1) / R parameter sets
    per step on every link direction, NVLS about one for any R: it pays from four ranks on (at two ranks both move one)."""
    have = bool(getattr(opt.symm_g, "has_multicast_support", False)) and int(getattr(opt.symm_g, "multicast_ptr", 0) or 0) != 0 \
        and int(getattr(opt.symm_p, "multicast_ptr", 0) or 0) != 0
    return have and world_size >= NVLS_MIN_RANKS


NVLS_MIN_RANKS = 4


def region_shard(off: int, n: int, rank: int, world_size: int):
    """Rank ``rank``'s part [a, b) of the range [off, off + n) of the flat buffers: R nearly equal pieces whose boundaries
    are multiples of 4 floats (16 bytes: float4 / TMA bulk / multimem granularity); n % 4 == 0."""
    q = n // 4
    return off + (q * rank // world_size) * 4, off + (q * (rank + 1) // world_size) * 4


class P2PExchange:
    """The data-parallel exchange fused into the optimizer sweep over NVLink peer memory, one contiguous RANGE of the flat
    buffers at a time (cb_clipped_adam_p2p / cb_clipped_adam_nvls).  ``region(off, n)``: device-side barrier (on every
    rank the gradient of the range is complete and its old parameter values are no longer read) -> ONE kernel that, for
    this rank's 1/R shard of the range, sums the gradient shards of all ranks straight out of their gradient buffers (in
    rank order), applies ClippedAdam and stores the new parameters into every rank's parameter buffer.  ``finish()``
    advances the step counter and closes the step with a second barrier (every shard delivered everywhere).

    The large weight blocks are exchanged from INSIDE the backward pass, each right behind its weight-gradient product on
    the bulk stream (optim.early_block_update), so the NVLink transfer of a block overlaps the rest of backward; the
    barriers and kernels are ordinary launches and are captured into the step graph with everything else.  Barriers of
    successive ranges use successive signal-pad channels (same order on every rank)."""

    def __init__(self, opt, world_size: int, rank: int, timeout_ms: int = 20000):
        self.world, self.rank = int(world_size), int(rank)
        self.timeout_ms = timeout_ms  # a rank that never arrives traps the waiting kernels instead of hanging the GPUs
        self.symm_g, self.symm_p = opt.symm_g, opt.symm_p
        f32 = torch.float32
        self._g_base = [self.symm_g.get_buffer(r, (opt.n,), f32).data_ptr() for r in range(self.world)]
        self._p_base = [self.symm_p.get_buffer(r, (opt.n,), f32).data_ptr() for r in range(self.world)]
        assert self._g_base[rank] == opt.flat_g.data_ptr(), "symmetric gradient buffer is not where flat_g lives"
        assert self._p_base[rank] == opt.flat_p.data_ptr(), "symmetric parameter buffer is not where flat_p lives"
        self.nvls = p2p_uses_nvls(opt, self.world)
        self.opt = opt
        self._channel = 0
        self.owned = set()  # (begin, end) element ranges whose moment buffers this rank keeps current

    def shard(self, off: int, n: int):
        return region_shard(off, n, self.rank, self.world)

    @torch.no_grad()
    def region(self, off: int, n: int):
        import ctypes

        from ._cabi import current_stream, lib

        if n <= 0:
            return
        assert off % 4 == 0 and n % 4 == 0, "exchange ranges must be 16-byte aligned"
        opt = self.opt
        c, self._channel = self._channel, self._channel + 1
        self.symm_g.barrier(channel=c, timeout_ms=self.timeout_ms)
        a, b = self.shard(off, n)
        if b <= a:
            return
        self.owned.add((a, b))
        at = lambda t: t.data_ptr() + 4 * a
        tab = (opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), opt.lr_table.numel(), opt.step_count.data_ptr(),
               float(opt.beta2), float(opt.eps), float(opt.clip), 0, current_stream())
        if self.nvls:  # in-switch reduction of the gradients, multicast store of the parameters
            lib().call("cb_clipped_adam_nvls", at(opt.flat_p), at(opt.flat_m), at(opt.flat_v), b - a,
                       int(self.symm_g.multicast_ptr) + 4 * a, int(self.symm_p.multicast_ptr) + 4 * a, *tab)
        else:
            grads = (ctypes.c_void_p * self.world)(*[g + 4 * a for g in self._g_base])
            peers = [p + 4 * a for r, p in enumerate(self._p_base) if r != self.rank]
            p_arr = (ctypes.c_void_p * max(1, len(peers)))(*peers)
            lib().call("cb_clipped_adam_p2p", at(opt.flat_p), at(opt.flat_m), at(opt.flat_v), b - a, ctypes.addressof(grads),
                       self.world, self.rank, ctypes.addressof(p_arr), len(peers), *tab)

    @torch.no_grad()
    def finish(self, opt=None):
        from ._cabi import current_stream, lib

        opt = self.opt
        at0 = lambda t: t.data_ptr()
        lib().call("cb_clipped_adam_range", at0(opt.flat_p), at0(opt.flat_g), at0(opt.flat_m), at0(opt.flat_v), 0,
                   opt.lr_table.data_ptr(), opt.beta1_table.data_ptr(), None, opt.lr_table.numel(), opt.step_count.data_ptr(),
                   float(opt.beta2), float(opt.eps), float(opt.clip), 1, current_stream())
        self.symm_p.barrier(channel=0, timeout_ms=self.timeout_ms)
        self._channel = 0


def exchange_regions(opt):
    """The ranges a whole-buffer exchange walks: every large weight block (laid out first), then the tail."""
    blocks = sorted({p._cb_block[0]: p._cb_block for p in opt.params
                     if getattr(p, "_cb_block", None) is not None and p._cb_block[0] < opt.tail_start}.values())
    out, cur = [], 0
    for o, rows, ld, cols in blocks:
        if o > cur:
            out.append((cur, o - cur))
        out.append((o, rows * ld))
        cur = o + rows * ld
    out.append((cur, opt.n - cur))
    return [(o, n) for o, n in out if n > 0]


@torch.no_grad()
def p2p_step(opt, world_size: int, rank: int, regions=None):
    """One whole exchange + sweep outside a backward pass (every range back to back; ``regions`` = [(offset, n)] tiling
    the buffer, default ``exchange_regions``): what tests/test_dp_single_gpu.py drives with virtual ranks."""
    if not opt.constant_learning_rate and opt.host_steps >= opt.total_steps:
        raise ValueError(f"Tried to step {opt.host_steps + 1} times. The specified number of total steps is {opt.total_steps}")
    ex = opt.exchange
    if ex is None:
        ex = opt.exchange = P2PExchange(opt, world_size, rank)
    for off, n in (exchange_regions(opt) if regions is None else regions):
        ex.region(off, n)
    ex.finish()
    opt.host_steps += 1


def exchange_microbench(opt, reps: int = 20):
    """Mean milliseconds of one whole exchange (all ranges back to back, barriers included) issued alone on an idle GPU:
    the cost the overlapped step hides.  Changes the parameters (real sweeps with the current gradients, step counter
    restored): call after the measurements that matter."""
    ex = opt.exchange
    if ex is None:
        return None
    step0 = opt.step_count.clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    regions = exchange_regions(opt)
    for i in range(reps + 3):
        if i == 3:
            ev[0].record()
        for off, n in regions:
            ex.region(off, n)
        ex.finish()
    ev[1].record()
    torch.cuda.synchronize()
    opt.step_count.copy_(step0)
    return ev[0].elapsed_time(ev[1]) / reps


@torch.no_grad()
def gather_optimizer_state(opt, world_size: int, rank: int):
    """Make the moment buffers complete on every rank after sharded stepping (call before saving a checkpoint)."""
    if world_size <= 1:
        return
    ex = getattr(opt, "exchange", None)
    if ex is not None:
        # fused exchange: every range of the buffers has its own 1/R shards; keep what this rank owns, sum over ranks
        mask = torch.zeros(opt.n, dtype=torch.float32, device=opt.flat_m.device)
        for a, b in ex.owned:
            mask[a:b] = 1.0
        for buf in (opt.flat_m, opt.flat_v):
            buf.mul_(mask)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
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


def _all_to_all_varlen(t: torch.Tensor, send_counts: List[int], recv_counts: List[int]) -> torch.Tensor:
    out = torch.empty(int(sum(recv_counts)), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out, t.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts))
    return out


@torch.no_grad()
def mckp_gene_sharded(piece, targets, n_genes: int, world_size: int, rank: int):
    """MultipleChoiceKnapsack over barcode-sharded posterior pieces: the knapsack of a gene needs the
    entries of ALL cells, so the pieces are re-keyed by gene with one all-to-all -- owner of gene g is rank g mod R
    (interleaved: highly and lowly expressed genes spread evenly) -- each rank runs ``estimation.mckp_device`` on its
    genes, and the per-(cell, gene) noise counts are gathered and merged by m.  ``piece``: this rank's DevicePosteriorCOO
    (cells of a contiguous barcode range); ``targets``: per-gene noise targets [n_genes] (every rank passes the same).
    Returns (seg_m int64, noise int32) of ALL segments, sorted by m, on every rank -- what ``mckp_device`` returns for the
    gathered COO, bit for bit (tests/test_multi_gpu.py)."""
    from .estimation import DevicePosteriorCOO, mckp_device

    if world_size <= 1:
        return mckp_device(piece, targets, n_genes)
    dev = piece.m.device

    def route(keys_m, *arrays):
        owner = (keys_m % n_genes) % world_size
        order = torch.argsort(owner, stable=True)  # stable: every destination keeps the (m, c) order of the piece
        send = torch.bincount(owner, minlength=world_size)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        send_l, recv_l = send.cpu().tolist(), recv.cpu().tolist()
        return [_all_to_all_varlen(a[order], send_l, recv_l) for a in (keys_m,) + arrays]

    m, c, lp = route(piece.m, piece.c, piece.log_prob)
    off_m, off_v = route(piece.off_m, piece.off_v)
    # pieces arrive in source-rank order = ascending barcode ranges: the concatenation is sorted by (m, c) already
    mine = DevicePosteriorCOO(m, c, lp, off_m, off_v, piece.n_cols)
    seg_m, noise = mckp_device(mine, targets, n_genes)
    seg_all, noise_all = all_gather_varlen(seg_m, world_size), all_gather_varlen(noise, world_size)
    order = torch.argsort(seg_all)
    return seg_all[order].contiguous(), noise_all[order].contiguous()


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
