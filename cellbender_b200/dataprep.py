"""Minibatch loader with the reference's interface and RNG behaviour (reference
cellbender/remove_background/data/dataprep.py:47-272), backed by a device-resident CSR.

What stays on the host, exactly as in the reference (so numpy's global RNG is consumed in the same order
and checkpoints of the loader state stay meaningful): the shuffled index list, the pointer, and the
``np.random.choice`` of surely-empty droplets (dataprep.py:117,171-178).  What moves to the GPU: the count
matrices (uploaded once) and the CSR -> dense gather + input transforms (``cb_csr_gather_rows``): the
reference's 27 ms/minibatch of scipy slicing + 34 MB H2D becomes a 2 KB index upload and one kernel.
"""
import logging
from typing import Callable, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch

from . import consts, ops
from .vae import EncoderInputs

logger = logging.getLogger("cellbender")


class MiniBatch:
    """One minibatch: row ids into the loader's device CSR (+ lazily gathered encoder inputs).
    Quacks like the dense tensor the reference yields where the training loop needs it (``size(0)``, ``shape``)."""

    def __init__(self, loader: "DataLoader", row_ids_host: np.ndarray):
        self.loader = loader
        self.row_ids_host = row_ids_host
        self.n = int(row_ids_host.size)
        self.row_ids = loader._upload_row_ids(row_ids_host)
        self._inputs: Optional[EncoderInputs] = None

    def size(self, dim: int = 0) -> int:
        return (self.n, self.loader.csr.n_genes)[dim]

    @property
    def shape(self):
        return (self.n, self.loader.csr.n_genes)

    def encoder_inputs(self, amb_unit: Optional[torch.Tensor] = None) -> EncoderInputs:
        out = ops.gather_rows(self.loader.csr, self.row_ids, amb_unit, want=("norm", "lognorm"),
                              out=self.loader._gather_buffers(self.n))
        return EncoderInputs(out["norm"], out["lognorm"], out["feats"], self.loader.csr.n_genes)

    def dense(self) -> torch.Tensor:
        """The dense fp32 count tensor the reference's loader yields (for tests / reference-style callers)."""
        out = ops.gather_rows(self.loader.csr, self.row_ids, None, want=("raw",))
        return out["raw"][:, : self.loader.csr.n_genes]


class DataLoader:
    """Loads a fraction of cell barcodes + unknowns and mixes in randomly sampled surely-empty droplets
    (reference dataprep.py:47-193; same constructor signature)."""

    def __init__(self, dataset: sp.csr_matrix, empty_drop_dataset: Optional[sp.csr_matrix],
                 batch_size: int = consts.DEFAULT_BATCH_SIZE, fraction_empties: float = consts.FRACTION_EMPTIES,
                 shuffle: bool = True, sort_by: Optional[Callable[[sp.csr_matrix], np.ndarray]] = None,
                 use_cuda: bool = True, device: Optional[str] = None):
        if not use_cuda:
            raise RuntimeError("cellbender_b200.DataLoader is device-resident (CUDA only)")
        if shuffle:
            assert sort_by is None, "Cannot sort_by and shuffle at the same time"
        self.sort_fn = sort_by
        self.dataset = dataset
        if self.sort_fn is not None:
            sort_order = np.argsort(self.sort_fn(self.dataset))
            self.ind_list = sort_order
            self.sort_order = sort_order.copy()
            self._unsort_dict = {i: val for i, val in enumerate(self.sort_order)}
        else:
            self.ind_list = np.arange(self.dataset.shape[0])
            self.sort_order = self.ind_list.copy()
            self._unsort_dict = {i: i for i in self.ind_list}
        self.empty_drop_dataset = empty_drop_dataset
        self.empty_ind_list = np.array([]) if empty_drop_dataset is None else np.arange(empty_drop_dataset.shape[0])
        self.batch_size = batch_size
        self.fraction_empties = fraction_empties
        self.cell_batch_size = int(batch_size * (1.0 - fraction_empties))
        self.shuffle = shuffle
        self.device = device or "cuda"
        self.use_cuda = True
        self._length = None
        # device-resident counts: analysed droplets first, then the surely-empty droplets
        stacked = dataset if (empty_drop_dataset is None or empty_drop_dataset.shape[0] == 0) \
            else sp.vstack([dataset, empty_drop_dataset], format="csr")
        self.csr = ops.DeviceCSR.from_scipy(stacked, self.device)
        self._n_cells_rows = dataset.shape[0]
        self._pinned = None
        self._dev_ids = None
        self._gbuf = None
        self._retired = []
        self._count_only = False
        self._max_batches: Optional[int] = None  # data parallel: every rank serves the same number of minibatches per epoch
        self._served = 0
        self._reset()

    # -- device plumbing ---------------------------------------------------------------------------------
    def _upload_row_ids(self, ids: np.ndarray) -> torch.Tensor:
        """2 KB of row ids, host -> device, through a small ring of pinned slots (the host may run several
        minibatches ahead of the GPU; a slot is reused only after its copy has completed)."""
        n = ids.size
        cap = int((self.batch_size + 8) * 1.5)
        if n > cap:
            return torch.as_tensor(ids.astype(np.int64), device=self.device)
        if self._pinned is None:
            self._ring = 8
            self._pinned = [torch.empty(cap, dtype=torch.int64).pin_memory() for _ in range(self._ring)]
            self._dev_ids = [torch.empty(cap, dtype=torch.int64, device=self.device) for _ in range(self._ring)]
            self._events = [None] * self._ring
            self._slot = 0
        k = self._slot
        self._slot = (k + 1) % self._ring
        if self._events[k] is not None:
            self._events[k].synchronize()
        self._pinned[k][:n].copy_(torch.from_numpy(ids.astype(np.int64, copy=False)))
        dst = self._dev_ids[k][:n]
        dst.copy_(self._pinned[k][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[k] = ev
        return dst

    def _gather_buffers(self, n: int) -> dict:
        if self._gbuf is None or self._gbuf["cap"] < n:
            G = self.csr.n_genes
            if self._gbuf is not None:
                self._retired.append(self._gbuf)  # captured step graphs refer to the old buffers by address: keep them alive
            self._gbuf = {"cap": n, "norm": ops.alloc_rows(n, G, self.device), "lognorm": ops.alloc_rows(n, G, self.device),
                          "feats": torch.empty((n, 4), device=self.device)}
        b = self._gbuf
        return {"norm": b["norm"][:n], "lognorm": b["lognorm"][:n], "feats": b["feats"][:n]}

    # -- reference interface ------------------------------------------------------------------------------
    @torch.no_grad()
    def unsort_inds(self, bcs):
        if self.sort_fn is None:
            return bcs
        return torch.tensor([self._unsort_dict[bc.item()] for bc in bcs], device="cpu")

    def _reset(self):
        if self.shuffle:
            np.random.shuffle(self.ind_list)
        self.ptr = 0
        self._served = 0

    def limit_batches(self, n: int):
        """Serve at most ``n`` minibatches per epoch (dp.agree_on_steps: ranks whose shard would yield one more stop
        early; the droplets left over are reshuffled into the next epoch like any others)."""
        self._max_batches = int(n)
        self._length = None

    def get_state(self):
        return {"ind_list": self.ind_list, "ptr": self.ptr}

    def set_state(self, ind_list: np.ndarray, ptr: int):
        self.ind_list = ind_list
        self.ptr = ptr
        assert self.ptr <= len(self.ind_list), (f"Problem setting dataloader state: pointer ({ptr}) is outside the "
                                                f"length of the ind_list ({len(ind_list)})")

    def reset_ptr(self):
        self.ptr = 0

    @property
    def length(self):
        if self._length is None:
            self._length = self._get_length()
        return self._length

    def _get_length(self):
        # like the reference, go through the loader once (this consumes numpy RNG: dataprep.py:140-145, run.py:762)
        self._count_only = True
        try:
            i = 0
            for _ in self:
                i += 1
        finally:
            self._count_only = False
        return i

    def __len__(self):
        return self.length

    def __iter__(self):
        return self

    def next_row_ids(self) -> np.ndarray:
        """Host part of __next__ (dataprep.py:153-193): indices of this minibatch in the stacked CSR."""
        remaining_cells = self.ind_list.size - self.ptr
        if self._max_batches is not None and self._served >= self._max_batches:
            self._reset()
            raise StopIteration()
        if remaining_cells < consts.SMALLEST_ALLOWED_BATCH:
            if remaining_cells > 0:
                logger.debug(f"Dropped last minibatch of {remaining_cells} cells")
            self._reset()
            raise StopIteration()
        next_ptr = min(self.ind_list.size, self.ptr + self.cell_batch_size)
        cell_inds = self.ind_list[self.ptr:next_ptr]
        n_empties = int(cell_inds.size * (self.fraction_empties / (1 - self.fraction_empties)))
        if self.empty_ind_list.size > 0:
            empty_inds = np.random.choice(self.empty_ind_list, size=n_empties, replace=True)
            ids = np.concatenate([cell_inds, empty_inds + self._n_cells_rows]) if empty_inds.size > 0 else cell_inds
        else:
            ids = cell_inds
        self.ptr = next_ptr
        self._served += 1
        return np.asarray(ids, dtype=np.int64)

    def __next__(self) -> MiniBatch:
        ids = self.next_row_ids()
        if self._count_only:
            return None
        return MiniBatch(self, ids)


def prep_sparse_data_for_training(dataset: sp.csr_matrix, empty_drop_dataset: sp.csr_matrix,
                                  training_fraction: float = consts.TRAINING_FRACTION,
                                  fraction_empties: float = consts.FRACTION_EMPTIES,
                                  batch_size: int = consts.DEFAULT_BATCH_SIZE, shuffle: bool = True,
                                  use_cuda: bool = True, device: Optional[str] = None) -> Tuple[DataLoader, DataLoader]:
    """Train / test loaders (reference dataprep.py:196-272; same numpy RNG consumption)."""
    training_mask = np.random.rand(dataset.shape[0]) < training_fraction
    training_indices = np.flatnonzero(training_mask)
    test_indices = np.flatnonzero(~training_mask)
    training_mask_empty = np.random.rand(empty_drop_dataset.shape[0]) < training_fraction
    training_indices_empty = np.flatnonzero(training_mask_empty)
    test_indices_empty = np.flatnonzero(~training_mask_empty)
    train_loader = DataLoader(dataset=dataset[training_indices, ...],
                              empty_drop_dataset=empty_drop_dataset[training_indices_empty, ...],
                              batch_size=batch_size, fraction_empties=fraction_empties, shuffle=shuffle, device=device)
    test_loader = DataLoader(dataset=dataset[test_indices, ...],
                             empty_drop_dataset=empty_drop_dataset[test_indices_empty, ...],
                             batch_size=batch_size, fraction_empties=fraction_empties, shuffle=shuffle, device=device)
    return train_loader, test_loader
