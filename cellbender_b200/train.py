"""Training loop with the reference's structure and criteria (reference cellbender/remove_background/train.py:
train_epoch :34-72, evaluate_epoch :75-104, run_training :108-300), driving the B200 step.

``SVI`` here is the pyro-free stand-in for ``pyro.infer.SVI(model.model, model.guide, scheduler,
TraceEnum_ELBO)`` (run.py:783-786): ``step(batch)`` = loss + backward + optimizer step, ``evaluate_loss(batch)``
= loss only.  Losses stay on the device; they are summed there and read once per epoch (the reference syncs every
step through ``loss.item()``).
"""
import logging
import sys
import time
from datetime import datetime
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _cabi
from .dataprep import DataLoader, MiniBatch
from .exceptions import ElboException, NanException
from .model import RemoveBackgroundPyroModel
from .optim import FlatClippedAdam

logger = logging.getLogger("cellbender")


class _StaticBatch:
    """Fixed-address stand-in for a MiniBatch (row ids live in a static device buffer) used under CUDA graphs."""

    def __init__(self, loader: DataLoader, n: int, device):
        self.loader = loader
        self.n = n
        self.row_ids = torch.zeros(n, dtype=torch.int64, device=device)

    def size(self, dim: int = 0) -> int:
        return (self.n, self.loader.csr.n_genes)[dim]

    encoder_inputs = MiniBatch.encoder_inputs


class SVI:
    """``step`` / ``evaluate_loss`` of pyro's SVI for this model.  With ``use_cuda_graph`` the whole training step
    (gather -> encoders -> sampling -> decoder -> fused likelihood fwd+bwd -> backward -> fused ClippedAdam) is
    captured once per minibatch size and replayed: one graph launch instead of ~10^3 kernel launches per step."""

    def __init__(self, model: RemoveBackgroundPyroModel, optim: FlatClippedAdam, use_cuda_graph: bool = False,
                 graph_warmup: int = 3):
        self.model = model
        self.optim = optim
        self.use_cuda_graph = use_cuda_graph
        self.graph_warmup = graph_warmup
        self._graphs = {}
        self._eager_steps = 0
        self._unit_grad = torch.ones((), dtype=torch.float32, device=model.device)  # d loss / d loss, never written to
        # data parallel (dp.attach): mode 'p2p' puts the exchange inside the step (optim.exchange), nothing to hook here;
        # mode 'sharded' runs the NCCL exchange eagerly between replays of the forward + backward graph
        self.after_backward = None  # hook(optim): gradient exchange that precedes a full optimizer sweep
        self.exchange_and_step = None  # hook(optim): exchange fused with / followed by the sharded optimizer sweep
        self.sync_offsets = None  # hook(): broadcast the encoder's first-forward offsets (called after the first step)
        self.sync_buffers = None  # hook(): average the BatchNorm running statistics over ranks (end of epoch / before eval)
        self.reduce_mean = None  # hook(float) -> float: mean over ranks (test ELBO for the failure criteria)

    def _ambient_unit_vector(self) -> torch.Tensor:
        """x_ambient / ||x_ambient||_2 of EncodeNonZLatents.forward (vae/encoder.py:266-270); the d_empty factor
        cancels in the normalisation."""
        with torch.no_grad():
            return self.model.ambient()[1]

    def loss_terms(self, batch, noise=None, row_keep_scale=None):
        # dropout rows, d_empty_loc, simplex transform of chi_ambient and the ambient-dot feature: beside the gather
        prologue = self.model.step_prologue(batch.n, row_keep_scale, batch.loader.csr, batch.row_ids)
        inp = batch.encoder_inputs(None)
        return self.model.elbo_terms(batch.loader.csr, batch.row_ids, inp, noise=noise, prologue=prologue)

    def _forward_backward(self, batch) -> torch.Tensor:
        from . import gemm

        self.optim.zero_grad()
        terms = self.loss_terms(batch)
        loss = terms["loss"]
        # single GPU: the ClippedAdam sweep of each large weight block is issued right behind its weight-gradient GEMM
        fuse = (self.after_backward is None and self.model.training and getattr(self.optim, "early_weight_updates", False))
        gemm.FUSED_OPT = self.optim if fuse else None
        gemm.UNIT_GRAD = self._unit_grad
        try:
            loss.backward(self._unit_grad)
        finally:
            gemm.FUSED_OPT = None
            gemm.UNIT_GRAD = None
        self.optim.collect_grads()
        return loss.detach()

    def _step_eager(self, batch) -> torch.Tensor:
        loss = self._forward_backward(batch)
        self._exchange_and_step()
        return loss

    def _exchange_and_step(self):
        if self.exchange_and_step is not None:
            self.exchange_and_step(self.optim)  # the exchange and the (sharded) sweep are one operation
            return
        if self.after_backward is not None:
            self.after_backward(self.optim)
        self.optim.step()

    def step(self, batch: MiniBatch) -> torch.Tensor:
        """One training step; returns the loss (-ELBO, summed over the minibatch) as a 0-dim device tensor."""
        if not self.use_cuda_graph or self.model.encoder["other"].offset is None or self._eager_steps < self.graph_warmup:
            self._eager_steps += 1  # first steps run eagerly (sets the encoder offsets, warms the allocator)
            first = self._eager_steps == 1
            loss = self._step_eager(batch)
            if first and self.sync_offsets is not None:
                self.sync_offsets()  # data parallel: every rank continues with rank 0's first-forward offsets
            return loss
        key = (id(batch.loader), batch.n, self.model.training)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(batch)
            self._graphs[key] = entry
        static, graph, loss, n_launch = entry
        static.row_ids.copy_(batch.row_ids)
        if not self.optim.constant_learning_rate and self.optim.host_steps >= self.optim.total_steps:
            raise ValueError(f"Tried to step {self.optim.host_steps + 1} times. The specified number of total steps is "
                             f"{self.optim.total_steps}")
        graph.replay()
        _cabi.lib().launch_count += n_launch  # kernels of ours inside the replayed graph
        if self.after_backward is not None:
            # data parallel, exchange outside the graph: all-reduce (NCCL) and the fused Adam run eagerly between replays
            self._exchange_and_step()
        else:
            self.optim.host_steps += 1
        return loss

    def _capture(self, batch: MiniBatch):
        static = _StaticBatch(batch.loader, batch.n, self.model.device)
        static.row_ids.copy_(batch.row_ids)
        # the captured step must not change training state twice: snapshot, run warm-up + capture, restore
        snap_p, snap_m, snap_v = self.optim.flat_p.clone(), self.optim.flat_m.clone(), self.optim.flat_v.clone()
        snap_step, host_steps = self.optim.step_count.clone(), self.optim.host_steps
        bn_state = {k: v.clone() for k, v in self.model.state_dict().items() if "running_" in k or "num_batches" in k}
        # single GPU: the whole step is one graph.  Data parallel: forward + backward + gradient collection only.
        body = self._step_eager if self.after_backward is None else self._forward_backward
        # warm-up and capture on a high-priority stream: the step's critical chain outranks the bulk stream (ops.side_stream)
        side = torch.cuda.Stream(priority=-1)
        side.wait_stream(torch.cuda.current_stream())
        # data parallel with the exchange inside the step: the warm-up run stays local (plain sweeps, no barriers) -- ranks
        # capture on their own (minibatch sizes differ between shards) and the state is restored below anyway
        self.optim.exchange_local = True
        try:
            with torch.cuda.stream(side):
                body(static)
        finally:
            self.optim.exchange_local = False
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        n0 = _cabi.lib().launch_count
        with torch.cuda.graph(graph, stream=side):
            loss = body(static)
        n_launch = _cabi.lib().launch_count - n0
        _cabi.lib().launch_count = n0  # capture records launches, it does not run them
        with torch.no_grad():
            self.optim.flat_p.copy_(snap_p), self.optim.flat_m.copy_(snap_m), self.optim.flat_v.copy_(snap_v)
            self.optim.step_count.copy_(snap_step)
            self.optim.host_steps = host_steps
            sd = self.model.state_dict()
            for k, v in bn_state.items():
                sd[k].copy_(v)
        return static, graph, loss, n_launch

    @torch.no_grad()
    def evaluate_loss(self, batch: MiniBatch) -> torch.Tensor:
        """-ELBO of a minibatch without a gradient step (pyro's ``SVI.evaluate_loss``; the test pass of train.py:75-104).
        With ``use_cuda_graph`` the forward pass is captured once per (loader, minibatch size) and replayed, like the
        training step: the test pass is ~300 launches per minibatch otherwise."""
        if (not self.use_cuda_graph or self.model.training or self.model.encoder["other"].offset is None
                or not isinstance(batch, MiniBatch)):
            return self.loss_terms(batch)["loss"].detach()
        key = (id(batch.loader), batch.n, "eval")
        entry = self._graphs.get(key)
        if entry is None:
            static = _StaticBatch(batch.loader, batch.n, self.model.device)
            static.row_ids.copy_(batch.row_ids)
            side = torch.cuda.Stream(priority=-1)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.loss_terms(static)  # warm-up outside capture
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = _cabi.lib().launch_count
            with torch.cuda.graph(graph, stream=side):
                loss = self.loss_terms(static)["loss"].detach()
            n_launch = _cabi.lib().launch_count - n0
            _cabi.lib().launch_count = n0
            entry = self._graphs[key] = (static, graph, loss, n_launch)
        static, graph, loss, n_launch = entry
        static.row_ids.copy_(batch.row_ids)
        graph.replay()
        _cabi.lib().launch_count += n_launch
        return loss.clone()


class LossReader:
    """Host-side reading of EVERY step's loss without stalling the device: ``push(loss)`` issues the device -> host copy of
    this step's loss into a pinned slot behind the step and returns the value of the PREVIOUS step, whose copy has long
    finished; ``flush()`` returns the last one.  The reference's ``svi.step`` returns ``loss.item()`` -- a device
    synchronisation per minibatch, during which the GPU idles while the host draws the next minibatch; with the read one
    step late the host logic of step t + 1 runs under step t.  (train_epoch needs no per-step value at all: it accumulates
    on the device and reads once per epoch.)"""

    def __init__(self, n_slots: int = 2):
        self._host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
        self._events = [torch.cuda.Event() for _ in range(n_slots)]
        self._pending = None
        self._i = 0

    def push(self, loss: torch.Tensor) -> Optional[float]:
        prev = self.flush()
        k = self._i % len(self._host)
        self._host[k].copy_(loss.detach().reshape(()), non_blocking=True)
        self._events[k].record()
        self._pending = k
        self._i += 1
        return prev

    def flush(self) -> Optional[float]:
        if self._pending is None:
            return None
        k, self._pending = self._pending, None
        self._events[k].synchronize()
        return float(self._host[k])


def is_scheduler(optim) -> bool:
    return isinstance(optim, FlatClippedAdam) and not optim.constant_learning_rate


def train_epoch(svi: SVI, train_loader: DataLoader) -> float:
    """-ELBO per droplet over one pass of the training loader (train.py:34-72)."""
    epoch_loss = torch.zeros((), device=svi.model.device)
    normalizer_train = 0.0
    for x_cell_batch in train_loader:
        epoch_loss += svi.step(x_cell_batch)  # optimizer step + scheduler step are fused (train.py:56-60)
        normalizer_train += x_cell_batch.size(0)
    svi.model.check_nan()
    return float(epoch_loss.item()) / normalizer_train


def evaluate_epoch(svi: SVI, test_loader: DataLoader) -> float:
    """Test-set loss (train.py:75-104)."""
    test_loss = torch.zeros((), device=svi.model.device)
    normalizer_test = 0.0
    for x_cell_batch in test_loader:
        test_loss += svi.evaluate_loss(x_cell_batch)
        normalizer_test += x_cell_batch.size(0)
    return float(test_loss.item()) / normalizer_test if normalizer_test > 0 else np.nan


def run_training(model: RemoveBackgroundPyroModel, svi: SVI, train_loader: DataLoader, test_loader: Optional[DataLoader],
                 epochs: int, test_freq: int = 10, final_elbo_fail_fraction: Optional[float] = None,
                 epoch_elbo_fail_fraction: Optional[float] = None, learning_rate: Optional[float] = None,
                 checkpoint_fn=None, checkpoint_freq: int = 10, debug: bool = False) -> Tuple[List[float], List[float]]:
    """Full course of training with periodic test-set evaluation and the reference's ELBO failure criteria
    (train.py:108-300).  ``checkpoint_fn(epoch)`` is called where the reference saves a checkpoint."""
    train_elbo: List[float] = []
    epoch_checkpoint_freq = 1000
    try:
        start_epoch = 1 if len(model.loss["train"]["epoch"]) == 0 else model.loss["train"]["epoch"][-1] + 1
        for epoch in range(start_epoch, epochs + 1):
            if epoch == start_epoch + 1:
                t = time.time()
            model.train()
            total_epoch_loss_train = train_epoch(svi, train_loader)
            train_elbo.append(-total_epoch_loss_train)
            last_learning_rate = svi.optim.get_last_lr()[0] if is_scheduler(svi.optim) else learning_rate
            model.loss["train"]["epoch"].append(epoch)
            model.loss["train"]["elbo"].append(-total_epoch_loss_train)
            model.loss["learning_rate"]["epoch"].append(epoch)
            model.loss["learning_rate"]["value"].append(last_learning_rate)
            if epoch == start_epoch + 1:
                time_per_epoch = time.time() - t
                logger.info("[epoch %03d]  average training loss: %.4f  (%.1f seconds per epoch)"
                            % (epoch, total_epoch_loss_train, time_per_epoch))
                epoch_checkpoint_freq = int(np.ceil(60 * checkpoint_freq / max(time_per_epoch, 1e-9)))
            else:
                logger.info("[epoch %03d]  average training loss: %.4f" % (epoch, total_epoch_loss_train))

            if svi.sync_buffers is not None:
                svi.sync_buffers()  # data parallel: one set of BatchNorm running statistics on every rank
            if (test_loader is not None) and (len(test_loader) > 0) and epoch % test_freq == 0:
                model.eval()
                total_epoch_loss_test = evaluate_epoch(svi, test_loader)
                if svi.reduce_mean is not None:  # the failure criteria below must fire on every rank or on none
                    total_epoch_loss_test = svi.reduce_mean(total_epoch_loss_test)
                model.loss["test"]["epoch"].append(epoch)
                model.loss["test"]["elbo"].append(-total_epoch_loss_test)
                logger.info("[epoch %03d] average test loss: %.4f" % (epoch, total_epoch_loss_test))
                if (epoch_elbo_fail_fraction is not None) and (len(model.loss["test"]["elbo"]) > 2):
                    current_diff = max(0.0, model.loss["test"]["elbo"][-2] - model.loss["test"]["elbo"][-1])
                    overall_diff = np.abs(model.loss["test"]["elbo"][-2] - model.loss["test"]["elbo"][0])
                    fractional_spike = current_diff / overall_diff
                    if fractional_spike > epoch_elbo_fail_fraction:
                        raise ElboException(
                            f"Training failed because test loss moved {current_diff:.2f} in the wrong direction, and "
                            f"that is {fractional_spike:.2f} of the total test ELBO change, > specified "
                            f"epoch_elbo_fail_fraction {epoch_elbo_fail_fraction:.2f}")

            if checkpoint_fn is not None and (((checkpoint_freq > 0) and (epoch % epoch_checkpoint_freq == 0))
                                              or (epoch == epochs)):
                checkpoint_fn(epoch)

        if final_elbo_fail_fraction is not None and len(model.loss["test"]["elbo"]) > 0:
            best_test_elbo = max(model.loss["test"]["elbo"])
            if model.loss["test"]["elbo"][-1] < best_test_elbo:
                final_best_diff = best_test_elbo - model.loss["test"]["elbo"][-1]
                initial_best_diff = best_test_elbo - model.loss["test"]["elbo"][0]
                if initial_best_diff == 0:
                    raise ElboException(
                        f"Training failed because there was no improvement from the initial test loss "
                        f"{model.loss['test']['elbo'][0]:.2f}. Final test loss was {model.loss['test']['elbo'][-1]}")
                elif (final_best_diff / initial_best_diff) > final_elbo_fail_fraction:
                    raise ElboException(
                        f"Training failed because final test loss {model.loss['test']['elbo'][-1]:.2f} is not "
                        f"sufficiently close to best test loss {best_test_elbo:.2f}, compared to the initial test loss "
                        f"{model.loss['test']['elbo'][0]:.2f}. Fractional difference is "
                        f"{final_best_diff / initial_best_diff:.2f}, which is > specified final_elbo_fail_fraction "
                        f"{final_elbo_fail_fraction:.2f}")

    except KeyboardInterrupt:
        logger.info("Inference procedure stopped by keyboard interrupt... will save a checkpoint.")
        if checkpoint_fn is not None:
            checkpoint_fn(-1)
    except ElboException as e:
        logger.info(e.message)
        raise e
    except NanException as nan:
        print(nan.message)
        logger.info(f"Inference procedure terminated early due to a NaN value in: {nan.param}\n\n"
                    f"The suggested fix is to reduce the learning rate by a factor of two.\n\n")
        sys.exit(1)

    logger.info(datetime.now().strftime("%Y-%m-%d %H:%M:%S"))
    if (final_elbo_fail_fraction is not None) and (len(model.loss["test"]["elbo"]) > 1):
        best_test_elbo = max(model.loss["test"]["elbo"])
        if -model.loss["test"]["elbo"][-1] >= -best_test_elbo * (1 + final_elbo_fail_fraction):
            raise ElboException(
                f"Training failed because final test loss ({-model.loss['test']['elbo'][-1]:.4f}) exceeds best test "
                f"loss ({-best_test_elbo:.4f}) by >= {100 * final_elbo_fail_fraction:.1f}%")
    torch.cuda.empty_cache()
    return train_elbo, model.loss["test"]["elbo"]
