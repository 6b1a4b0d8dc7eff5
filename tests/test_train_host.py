"""Host logic of run_training (reference train.py:108-300) with scripted losses and no GPU: the epoch / test-pass
bookkeeping, both test-ELBO failure criteria, the NaN exit, checkpoint calls and the keyboard-interrupt path."""
import pytest
import torch

from cellbender_b200 import train
from cellbender_b200.exceptions import ElboException, NanException


class _Batch:
    def size(self, dim=0):
        return 10


class _Loader:
    def __init__(self, n):
        self.n = n

    def __iter__(self):
        return iter([_Batch() for _ in range(self.n)])

    def __len__(self):
        return self.n


class _Model:
    device = "cpu"

    def __init__(self, nan_at_epoch=None):
        self.loss = {"train": {"epoch": [], "elbo": []}, "test": {"epoch": [], "elbo": []},
                     "learning_rate": {"epoch": [], "value": []}}
        self.nan_at_epoch, self.epochs_checked, self.mode = nan_at_epoch, 0, None

    def train(self):
        self.mode = "train"

    def eval(self):
        self.mode = "eval"

    def check_nan(self):
        self.epochs_checked += 1
        if self.nan_at_epoch == self.epochs_checked:
            raise NanException(param="mu")


class _Svi:
    """Per-droplet losses: training 50 - epoch, test as scripted (one value per test pass)."""
    sync_buffers = reduce_mean = None
    optim = object()

    def __init__(self, model, test_losses, interrupt_at_step=None):
        self.model, self.test_losses, self.tests_done, self.steps = model, list(test_losses), 0, 0
        self.interrupt_at_step = interrupt_at_step

    def step(self, batch):
        self.steps += 1
        if self.interrupt_at_step == self.steps:
            raise KeyboardInterrupt()
        assert self.model.mode == "train"
        epoch = (self.steps - 1) // 3 + 1
        return torch.tensor(10.0 * (50.0 - epoch))

    def evaluate_loss(self, batch):
        assert self.model.mode == "eval"
        return torch.tensor(10.0 * self.test_losses[self.tests_done // 2])  # two test minibatches per pass

    def count(self):
        self.tests_done += 1


def _run(test_losses, epochs, nan_at_epoch=None, interrupt_at_step=None, **kw):
    model = _Model(nan_at_epoch)
    svi = _Svi(model, test_losses, interrupt_at_step)
    inner = svi.evaluate_loss

    def counted(batch):
        out = inner(batch)
        svi.count()
        return out

    svi.evaluate_loss = counted
    saved = []
    out = train.run_training(model, svi, _Loader(3), _Loader(2), epochs=epochs, test_freq=1, learning_rate=1e-4,
                             checkpoint_fn=saved.append, **kw)
    return model, svi, saved, out


def test_bookkeeping_of_a_healthy_run():
    model, svi, saved, (train_elbo, test_elbo) = _run([40.0, 30.0, 25.0, 24.0], epochs=4, epoch_elbo_fail_fraction=0.5,
                                                      final_elbo_fail_fraction=0.5)
    assert train_elbo == [-49.0, -48.0, -47.0, -46.0] == model.loss["train"]["elbo"]
    assert test_elbo == [-40.0, -30.0, -25.0, -24.0] and model.loss["test"]["epoch"] == [1, 2, 3, 4]
    assert model.loss["learning_rate"]["value"] == [1e-4] * 4 and svi.steps == 12
    assert saved == [4]  # the reference saves at the last epoch (train.py:259-262)


def test_epoch_spike_criterion_fires():
    """train.py:211-226: the test loss moving the wrong way by more than the given fraction of the total change so far."""
    with pytest.raises(ElboException, match="wrong direction"):
        _run([40.0, 30.0, 38.0, 20.0], epochs=4, epoch_elbo_fail_fraction=0.5)  # 8 back out of 10 gained = 0.8 > 0.5
    _run([40.0, 30.0, 33.0, 20.0], epochs=4, epoch_elbo_fail_fraction=0.5)  # 0.3 <= 0.5: tolerated


def test_final_criterion_fires_when_the_run_ends_far_from_its_best():
    """train.py:264-285."""
    with pytest.raises(ElboException, match="not sufficiently close to best test loss"):
        _run([40.0, 20.0, 35.0], epochs=3, final_elbo_fail_fraction=0.5)  # ends 15 of 20 away from the best
    with pytest.raises(ElboException, match="no improvement from the initial test loss"):
        _run([40.0, 45.0, 50.0], epochs=3, final_elbo_fail_fraction=0.5)
    _run([40.0, 20.0, 22.0], epochs=3, final_elbo_fail_fraction=0.5)


def test_nan_terminates_the_process_like_the_reference():
    """train.py:277-285: a NanException ends the run with exit code 1."""
    with pytest.raises(SystemExit) as e:
        _run([40.0, 30.0], epochs=2, nan_at_epoch=2)
    assert e.value.code == 1


def test_keyboard_interrupt_saves_a_checkpoint_and_returns():
    model, svi, saved, (train_elbo, test_elbo) = _run([40.0, 30.0, 25.0], epochs=3, interrupt_at_step=5)
    assert saved == [-1] and len(train_elbo) == 1 and test_elbo == [-40.0]
