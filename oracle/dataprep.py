"""ORACLE (test infrastructure, NOT the product path): host restatement of the reference minibatch loader
(reference cellbender/remove_background/data/dataprep.py: DataLoader.__next__ :153-193, sparse_collate :275-292):
scipy fancy row indexing of the cell matrix and of the surely-empty matrix, vstack, todense, float32.
Pinned by tests/golden/dataloader.npz (generated from the unmodified reference)."""
import numpy as np
import scipy.sparse as sp
import torch

SMALLEST_ALLOWED_BATCH = 4  # consts.py:111


class HostLoader:
    def __init__(self, dataset: sp.csr_matrix, empties, batch_size=128, fraction_empties=0.2, shuffle=True):
        self.dataset, self.empties = dataset, empties
        self.ind_list = np.arange(dataset.shape[0])
        self.empty_ind_list = np.array([]) if empties is None else np.arange(empties.shape[0])
        self.fraction_empties = fraction_empties
        self.cell_batch_size = int(batch_size * (1.0 - fraction_empties))
        self.shuffle = shuffle
        self._reset()

    def _reset(self):
        if self.shuffle:
            np.random.shuffle(self.ind_list)
        self.ptr = 0

    def __iter__(self):
        return self

    def __len__(self):
        n = 0
        for _ in self:
            n += 1
        return n

    def __next__(self) -> torch.Tensor:
        remaining = self.ind_list.size - self.ptr
        if remaining < SMALLEST_ALLOWED_BATCH:
            self._reset()
            raise StopIteration()
        nxt = min(self.ind_list.size, self.ptr + self.cell_batch_size)
        cell_inds = self.ind_list[self.ptr:nxt]
        n_empties = int(cell_inds.size * (self.fraction_empties / (1 - self.fraction_empties)))
        mats = [self.dataset[cell_inds, :]]
        if self.empty_ind_list.size > 0:
            empty_inds = np.random.choice(self.empty_ind_list, size=n_empties, replace=True)
            if empty_inds.size > 0:
                mats.append(self.empties[empty_inds, :])
        self.ptr = nxt
        return torch.from_numpy(np.array(sp.vstack(mats, format="csr").todense(), dtype=np.float32))
