"""Host logic of the flat optimizer buffers (no GPU needed): parameter views alias ONE buffer, large matrices are laid
out first with 128-byte-aligned padded rows, the tail is contiguous, and the OneCycle tables are torch's own."""
import torch
import torch.nn as nn

from cellbender_b200 import optim


def _toy_model(big_numel):
    m = nn.ModuleDict({
        "a": nn.Linear(37, 5),                       # small, 37 % 4 != 0: rows padded to 40 floats (TMA needs 16-byte rows)
        "a2": nn.Linear(36, 5),                      # small, 16-byte rows already: no padding
        "big1": nn.Linear(1030, big_numel // 1030 + 1),  # >= BIG_BLOCK_NUMEL once multiplied out (see below)
        "mid": nn.Linear(258, 300),                  # >= 65536 elements: padded rows (258 -> 288), but not "big"
        "b": nn.BatchNorm1d(11),
        "big2": nn.Linear(big_numel // 512 + 1, 512),
    })
    return m


def test_flat_layout_big_blocks_first_and_views_alias(monkeypatch):
    monkeypatch.setattr(optim, "BIG_BLOCK_NUMEL", 1 << 17)
    model = _toy_model(1 << 17)
    model["mid"].weight._cb_lead = 28  # a Linear(cat(feat4, x)) weight: its wide block starts on a 128-byte line
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    opt = optim.FlatClippedAdam(model.parameters(), learning_rate=1e-3, total_steps=7)
    names = [k for k, _ in model.named_parameters()]
    big = [i for i, p in enumerate(opt.params) if p.dim() == 2 and p.numel() >= (1 << 17)]
    assert len(big) == 2
    # values preserved, parameters are views of the flat buffer, gradient views line up with them
    for (k, p), o, ld in zip(model.named_parameters(), opt.offsets, opt.lds):
        assert torch.equal(p.detach(), before[k]), k
        lead = getattr(p, "_cb_lead", 0) if ld is not None else 0
        assert p.data_ptr() == opt.flat_p.data_ptr() + 4 * (o + lead), k
        assert p._cb_grad_view.data_ptr() == opt.flat_g.data_ptr() + 4 * (o + lead) and p._cb_grad_view.shape == p.shape
        assert o % 4 == 0  # 16-byte aligned start
        if p.dim() == 2 and p.numel() >= (1 << 16):
            assert ld == (lead, (lead + p.shape[1] + 31) // 32 * 32) and p.stride(0) == ld[1] and o % 32 == 0
            assert p._cb_block == (o, p.shape[0], ld[1], p.shape[1])
        elif p.dim() == 2 and p.shape[1] % 4 != 0:  # small layer with rows TMA could not address: 16-byte row pitch
            assert ld == (0, (p.shape[1] + 3) // 4 * 4) and p.stride(0) == ld[1] and o % 32 == 0
        else:
            assert ld is None and not hasattr(p, "_cb_block")
    # big blocks occupy [0, tail_start) back to back; every other tensor lives in the tail
    spans = sorted((opt.offsets[i], opt.offsets[i] + opt.params[i].shape[0] * opt.lds[i][1]) for i in big)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == opt.tail_start
    assert all(opt.offsets[i] >= opt.tail_start for i in range(len(opt.params)) if i not in big)
    assert opt.n % 64 == 0 and opt.flat_p.numel() == opt.n == opt.flat_m.numel() == opt.flat_v.numel()
    # writing through a parameter is visible in the flat buffer (padding columns stay zero)
    i = names.index("mid.weight")
    with torch.no_grad():
        opt.params[i].fill_(2.0)
    blk = opt.flat_p[opt.offsets[i]: opt.offsets[i] + 300 * 288].view(300, 288)
    assert float(blk[:, 28:286].min()) == 2.0 and float(blk[:, 286:].abs().max()) == 0.0 and float(blk[:, :28].abs().max()) == 0.0
    # the block behind the 4 feature columns is line-aligned (relative to the buffer, which CUDA allocates 512-byte aligned)
    assert (opt.params[i].data_ptr() - opt.flat_p.data_ptr() + 4 * 4) % 128 == 0


def test_one_cycle_tables_are_torchs_schedule():
    """run.py:604-619 / tests/test_train.py:87-108 of the reference: OneCycleLR(max_lr = 10 lr), cycling beta1."""
    lrs, b1s = optim.one_cycle_tables(1e-4, 40)
    assert len(lrs) == len(b1s) == 40
    assert abs(lrs[0] - 1e-3 / 25) < 1e-12 and abs(max(lrs) - 1e-3) < 1e-9 and abs(lrs[-1] - 1e-3 / 25 / 1e4) < 1e-12
    assert abs(b1s[0] - 0.95) < 1e-9 and abs(min(b1s) - 0.85) < 1e-6 and abs(b1s[-1] - 0.95) < 1e-6
    assert lrs.index(max(lrs)) == b1s.index(min(b1s))  # momentum is lowest where the learning rate peaks
