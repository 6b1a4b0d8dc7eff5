import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def fp32_gemm(monkeypatch):
    """Substitute the product's only GEMM entry (ops.gemm_tf32: tcgen05 kind::tf32) by a full-fp32 library product with
    the same signature, for the parity tests that separate 'the ELBO algebra is right to 1e-5' from 'TF32 rounds the
    operands of the large contractions'.  Test infrastructure only: the product has no such switch."""
    import torch

    from cellbender_b200 import ops

    def gemm_fp32(a, b, M, N, K, a_mn=False, b_mn=False, out=None, bias=None, accumulate=False, split_k=None, a2=None,
                  b2=None, K2=0, tile_cfg=0):
        def operand(t, rows, k, mn):
            return t[:k, :rows].t() if mn else t[:rows, :k]

        c = operand(a, M, K, a_mn) @ operand(b, N, K, b_mn).t()
        if a2 is not None:
            c = c + operand(a2, M, K2, a_mn) @ operand(b2, N, K2, b_mn).t()
        if bias is not None:
            c = c + bias
        if out is None:
            out = torch.empty((M, ops.round_up(N, 4)), dtype=torch.float32, device=a.device)[:, :N]
        if accumulate:
            out[:M, :N] += c
        else:
            out[:M, :N] = c
        return out

    monkeypatch.setattr(ops, "gemm_tf32", gemm_fp32)
    yield
