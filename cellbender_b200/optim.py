"""Optimizer: pyro's ClippedAdam driven by torch's OneCycleLR (reference run.py:593-629), fused over one flat
parameter buffer (csrc/adam.cu).

Reference semantics kept: one optimizer step per minibatch followed by one scheduler step (train.py:56-60);
``OneCycleLR(max_lr = 10 lr, total_steps = n_batches * epochs)`` with torch's defaults (pct_start 0.3, cosine
annealing, div_factor 25, final_div_factor 1e4, cycle_momentum: beta1 0.95 -> 0.85 -> 0.95); elementwise gradient
clamp at +-10 (pyro ClippedAdam ``clip_norm``), beta2 0.999, eps 1e-8, no weight decay, lrd 1; stepping beyond
``total_steps`` raises like torch's scheduler does (pinned by the reference's tests/test_train.py:87-108).
``constant_learning_rate=True`` gives plain ClippedAdam at ``lr`` with beta1 0.9 (run.py:621-627).
"""
from typing import Iterable, List, Optional

import torch
import torch.nn as nn

from . import ops


BIG_BLOCK_NUMEL = 1 << 22  # the G x 512 matrices (17 M elements each at the PBMC8k shape)


def one_cycle_tables(learning_rate: float, total_steps: int):
    """Per-step (lr, beta1) of torch.optim.lr_scheduler.OneCycleLR wrapped around an Adam-family optimizer."""
    dummy = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=learning_rate, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.OneCycleLR(dummy, max_lr=learning_rate * 10, total_steps=total_steps)
    lrs, b1s = [], []
    for k in range(total_steps):
        lrs.append(dummy.param_groups[0]["lr"])
        b1s.append(dummy.param_groups[0]["betas"][0])
        dummy.step()
        if k < total_steps - 1:
            sched.step()
    return lrs, b1s


class FlatClippedAdam:
    """All trainable tensors re-homed into ONE contiguous fp32 buffer (the nn.Parameters become views, names and
    shapes unchanged), with flat gradient / first-moment / second-moment buffers beside it."""

    def __init__(self, params: Iterable[nn.Parameter], learning_rate: float, total_steps: Optional[int],
                 constant_learning_rate: bool = False, clip_norm: float = 10.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 symmetric_group=None):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no parameters to optimise"
        dev = self.params[0].device
        self.device = dev
        # each tensor starts on a 16-byte boundary; large 2-D weights additionally get 128-byte-aligned rows (a row pitch
        # that is a multiple of 32 floats; G = 33 538 is not even a multiple of 4) -- the nn.Parameter then is a strided
        # view [rows, cols] of a [rows, ld] block (gemm.padded_layout), the padding columns stay zero forever.
        from .gemm import needs_padded_layout, padded_layout

        def padded_ld(p):
            return padded_layout(p) if needs_padded_layout(p) else None

        # layout: the few very large matrices first (each one is a unit of the piecewise optimizer sweep and, data
        # parallel, of the overlapped gradient reduction), everything else behind them in ONE contiguous tail
        big = [p.dim() == 2 and p.numel() >= BIG_BLOCK_NUMEL for p in self.params]
        order = [i for i, b in enumerate(big) if b] + [i for i, b in enumerate(big) if not b]
        self.offsets, self.lds, off = [0] * len(self.params), [None] * len(self.params), 0
        self.tail_start = 0
        for i in order:
            p = self.params[i]
            ld = padded_ld(p)  # None or (lead padding, row pitch)
            n = p.numel() if ld is None else p.shape[0] * ld[1]
            if ld is not None:
                off = (off + 31) // 32 * 32  # padded blocks start on 128-byte lines
            self.offsets[i] = off
            self.lds[i] = ld
            off += (n + 3) // 4 * 4
            if big[i]:
                self.tail_start = off
        self.n = (off + 63) // 64 * 64  # a multiple of 4 x (1, 2, 4, 8, 16) ranks: equal 16-byte-aligned shards (dp.py)
        off = self.n
        # data parallel with the exchange fused into the optimizer sweep (dp.p2p_step): the parameter and gradient
        # buffers are symmetric-memory allocations, mapped into every rank's address space over NVLink
        self.symm_p = self.symm_g = None
        if symmetric_group is not None:
            import torch.distributed._symmetric_memory as symm_mem

            self.flat_p = symm_mem.empty(off, dtype=torch.float32, device=dev).zero_()
            self.flat_g = symm_mem.empty(off, dtype=torch.float32, device=dev).zero_()
            self.symm_p = symm_mem.rendezvous(self.flat_p, symmetric_group)
            self.symm_g = symm_mem.rendezvous(self.flat_g, symmetric_group)
        else:
            self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
            self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o, ld in zip(self.params, self.offsets, self.lds):
                if ld is None:
                    view = self.flat_p[o:o + p.numel()].view(p.shape)
                    gview = self.flat_g[o:o + p.numel()].view(p.shape)
                else:
                    rows, cols = p.shape
                    lead, ld = ld
                    view = self.flat_p[o:o + rows * ld].view(rows, ld)[:, lead:lead + cols]
                    gview = self.flat_g[o:o + rows * ld].view(rows, ld)[:, lead:lead + cols]
                view.copy_(p.data)
                p.data = view
                p._cb_grad_view = gview  # GEMM epilogues write weight gradients here directly (gemm.py)
                if ld is not None:
                    p._cb_block = (o, p.shape[0], ld, p.shape[1])  # [rows, ld] block of the flat buffers
                self.grad_views.append(gview)
        self.learning_rate = learning_rate
        self._beta1_constant = betas[0]
        self._constant_requested = constant_learning_rate
        self.beta2, self.eps, self.clip = betas[1], eps, clip_norm
        self.set_total_steps(total_steps)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)  # optimizer steps taken (device side)
        self.host_steps = 0
        # Default on a single GPU: the ClippedAdam sweep of each large weight block is issued as soon as that block's
        # gradient GEMM has finished, on the same (side) stream, so that the HBM-bound sweep overlaps the rest of the
        # backward pass (which is latency / L2-bound); the sweep at the end of the step covers what is left.
        self.early_weight_updates = True
        # data parallel with sharded sweeps (dp.attach): callable that completes flat_m / flat_v on every rank
        self.gather_state = None
        # data parallel over NVLink peer memory (dp.attach, mode 'p2p'): object whose ``region(offset, n)`` performs the
        # exchange + sharded sweep of one contiguous range of the flat buffers and whose ``finish()`` is the closing barrier;
        # ``exchange_local`` = True while a step graph is warmed up before capture (no communication: ranks capture alone)
        self.exchange = None
        self.exchange_local = False
        self._fused_done: List[tuple] = []

    def set_total_steps(self, total_steps: Optional[int]):
        """(Re)build the per-step lr / beta1 tables of OneCycleLR for ``total_steps`` optimizer steps."""
        self.total_steps = total_steps
        # epochs == 0 ("will only initialize the model", run.py:789): nothing to schedule
        self.constant_learning_rate = self._constant_requested or total_steps is None or total_steps <= 0
        if self.constant_learning_rate:
            lrs, b1s = [self.learning_rate], [self._beta1_constant]
        else:
            lrs, b1s = one_cycle_tables(self.learning_rate, total_steps)
        self._lrs = lrs
        self.lr_table = torch.tensor(lrs, dtype=torch.float32, device=self.device)
        self.beta1_table = torch.tensor(b1s, dtype=torch.float32, device=self.device)
        # bias-corrected step size of step t (1-based), lr_t sqrt(1 - b2^t) / (1 - b1_t^t), from the fp32 table values in
        # float64 -- exactly what the kernels would evaluate; with a constant learning rate t is unbounded: no table
        self.step_table = None
        if not self.constant_learning_rate and hasattr(self, "beta2"):
            self._build_step_table()

    def _build_step_table(self):
        import numpy as np

        lr = self.lr_table.double().cpu().numpy()
        b1 = self.beta1_table.double().cpu().numpy()
        t = np.arange(1, lr.size + 1, dtype=np.float64)
        step = lr * np.sqrt(1.0 - np.float64(np.float32(self.beta2)) ** t) / (1.0 - b1 ** t)
        self.step_table = torch.tensor(step, dtype=torch.float32, device=self.device)

    def zero_grad(self):
        self._fused_done = []
        for p in self.params:
            p.grad = None
            p._cb_grad_direct = False

    def collect_grads(self):
        """Gather the autograd-produced gradients into the flat buffer (missing gradient = 0)."""
        src, dst = [], []
        for p, gv in zip(self.params, self.grad_views):
            if p.grad is None:
                if not getattr(p, "_cb_grad_direct", False):  # else: a GEMM epilogue already wrote this gradient
                    gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                src.append(p.grad)
                dst.append(gv)
        if src:
            torch._foreach_copy_(dst, src)

    def early_block_update(self, w: nn.Parameter):
        """ClippedAdam sweep of one [rows, ld] weight block of the flat buffers, on the current stream, without advancing
        the step counter.  The caller guarantees the block's gradient is complete and its old values are no longer read."""
        from ._cabi import current_stream, lib, ptr

        o, rows, ld, cols = w._cb_block
        self._fused_done.append((o, rows, ld, cols, 0))
        if self.exchange is not None and not self.exchange_local:
            self.exchange.region(o, rows * ld)  # all ranks: gradient block complete, old weights no longer read
            return
        at = lambda t: t.data_ptr() + 4 * o
        lib().call("cb_clipped_adam_range", at(self.flat_p), at(self.flat_g), at(self.flat_m), at(self.flat_v), rows * ld,
                   self.lr_table.data_ptr(), self.beta1_table.data_ptr(), ptr(self.step_table), self.lr_table.numel(),
                   self.step_count.data_ptr(), float(self.beta2), float(self.eps), float(self.clip), 0, current_stream())

    def step(self, grad_scale: float = 1.0):
        """One ClippedAdam step over the flat buffers + one scheduler step."""
        if not self.constant_learning_rate and self.host_steps >= self.total_steps:
            raise ValueError(f"Tried to step {self.host_steps + 1} times. The specified number of total steps is "
                             f"{self.total_steps}")
        from . import gemm

        if self.exchange is not None and not self.exchange_local:
            # data parallel: join the bulk stream (the large blocks' exchanges, issued from inside backward), then exchange
            # the ranges in between (everything but the large blocks), advance the step counter and close the step with the
            # barrier.  The join comes FIRST so that on every rank the barriers form one dependency chain in the same
            # order: whatever order the hardware serialises independent graph branches in, a spinning barrier can then
            # never sit in front of a kernel that an earlier barrier of a peer is waiting for.
            gemm.join_bulk()
            cur = 0
            for o, rows, ld, cols, off in sorted(self._fused_done):
                assert o >= cur, "fused blocks overlap"
                self.exchange.region(cur, o - cur)
                cur = o + rows * ld
            self.exchange.region(cur, self.n - cur)
            self.exchange.finish()
            self._fused_done = []
            self.host_steps += 1
            return
        if self._fused_done:
            # blocks already swept behind their weight-gradient GEMM (early_block_update): sweep the ranges between them;
            # nothing here depends on the bulk stream, which is joined only before the step counter advances
            from ._cabi import current_stream, lib, ptr

            assert grad_scale == 1.0, "piecewise weight updates do not support gradient rescaling"
            L, st = lib(), current_stream()
            tab = (self.lr_table.data_ptr(), self.beta1_table.data_ptr(), ptr(self.step_table), self.lr_table.numel(),
                   self.step_count.data_ptr(), float(self.beta2), float(self.eps), float(self.clip))
            at = lambda t, off: t.data_ptr() + 4 * off
            cur = 0
            for o, rows, ld, cols, off in sorted(self._fused_done):
                assert o >= cur, "fused blocks overlap"
                L.call("cb_clipped_adam_range", at(self.flat_p, cur), at(self.flat_g, cur), at(self.flat_m, cur),
                       at(self.flat_v, cur), o - cur, *tab, 0, st)
                cur = o + rows * ld
            L.call("cb_clipped_adam_range", at(self.flat_p, cur), at(self.flat_g, cur), at(self.flat_m, cur), at(self.flat_v, cur),
                   self.n - cur, *tab, 0, st)
            gemm.join_bulk()  # the early block sweeps read the step counter: advance it only behind them
            L.call("cb_clipped_adam_range", at(self.flat_p, 0), at(self.flat_g, 0), at(self.flat_m, 0), at(self.flat_v, 0),
                   0, *tab, 1, current_stream())
            self._fused_done = []
        else:
            gemm.join_bulk()
            ops.clipped_adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr_table, self.beta1_table,
                                  self.step_count, beta2=self.beta2, eps=self.eps, clip=self.clip, grad_scale=grad_scale,
                                  step_table=self.step_table)
        self.host_steps += 1

    def get_last_lr(self) -> List[float]:
        """LR that the NEXT optimizer step will use (what OneCycleLR.get_last_lr() reports after its step)."""
        if self.constant_learning_rate:
            return [self.learning_rate]
        return [self._lrs[min(self.host_steps, len(self._lrs) - 1)]]

    def state_dict(self):
        if self.gather_state is not None:
            self.gather_state()  # sharded data-parallel sweeps keep only this rank's slice of the moments current
        return {"flat_m": self.flat_m, "flat_v": self.flat_v, "step": self.step_count, "host_steps": self.host_steps,
                "total_steps": self.total_steps, "learning_rate": self.learning_rate}

    def load_state_dict(self, sd):
        self.flat_m.copy_(sd["flat_m"])
        self.flat_v.copy_(sd["flat_v"])
        self.step_count.copy_(sd["step"])
        self.host_steps = int(sd["host_steps"])


def get_optimizer(model, n_batches: int, batch_size: int, epochs: int, learning_rate: float,
                  constant_learning_rate: bool, total_epochs_for_testing_only: Optional[int] = None,
                  symmetric_group=None) -> FlatClippedAdam:
    """Reference run.py:593-629 (same arguments, plus the model whose parameters are optimised; ``symmetric_group``:
    process group over which the flat parameter / gradient buffers become NVLink peer memory, see dp.py)."""
    total_steps = n_batches * (total_epochs_for_testing_only if total_epochs_for_testing_only is not None else epochs)
    return FlatClippedAdam(model.parameters(), learning_rate, total_steps, constant_learning_rate=constant_learning_rate,
                           symmetric_group=symmetric_group)
