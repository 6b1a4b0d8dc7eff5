"""CPU tests of the host logic around the output-assembly kernels: value-type promotion, the dp exchange-mode
selection, and that nothing falls back to the CPU when no CUDA device is present."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from cellbender_b200 import dp, sparse_utils as su


def test_value_dtype_promotion_follows_numpy_result_type():
    i32, i64, f32, f64 = (np.dtype(t) for t in (np.int32, np.int64, np.float32, np.float64))
    assert su._value_dtype(i32) == i32 and su._value_dtype(i32, i32) == i32
    assert su._value_dtype(i32, i64) == i64 and su._value_dtype(np.dtype(np.uint32), i32) == i64
    assert su._value_dtype(np.dtype(np.uint32)) == i64  # does not fit int32
    assert su._value_dtype(np.dtype(np.uint16)) == i32 and su._value_dtype(np.dtype(bool)) == i32
    assert su._value_dtype(i32, f32) == f64  # numpy: int32 + float32 -> float64, what scipy's binop returns
    assert su._value_dtype(np.dtype(np.int16), f32) == f32 and su._value_dtype(np.dtype(np.float16)) == f32
    assert su._value_dtype(f32, f64) == f64
    with pytest.raises(TypeError):
        su._value_dtype(np.dtype(np.complex64))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback():
    a = sp.csr_matrix(np.eye(3, dtype=np.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        su.combine_csr(a, a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        su.overwrite_matrix_with_columns_from_another(a, a, [0])
    with pytest.raises(RuntimeError):
        su.dense_to_sparse_op_torch(torch.zeros(2, 2))
    with pytest.raises(ValueError):
        su.csr_set_rows_to_zero(np.zeros((2, 2)), [0])  # not a sparse matrix: the reference's error


def test_dp_mode_selection():
    """dp.attach picks the exchange once, from what the optimizer's buffers are (no environment switch): 'p2p' needs
    symmetric-memory buffers, otherwise the NCCL path; unknown names are refused."""
    assert dp.p2p_group() is None  # torch.distributed is not initialised here: no peer-memory group

    class Opt:
        symm_p = None
        n = 64

    class Svi:
        optim = Opt()
        after_backward = exchange_and_step = None

    assert dp.attach(Svi(), 1).after_backward is None  # world size 1: untouched


def test_p2p_uses_nvls_policy():
    class H:
        def __init__(self, ok):
            self.has_multicast_support, self.multicast_ptr = ok, (4096 if ok else 0)

    class Opt:
        def __init__(self, ok):
            self.symm_g, self.symm_p = H(ok), H(ok)

    assert dp.p2p_uses_nvls(Opt(True), 2) is False  # two ranks: plain peer loads / stores move the same bytes
    assert dp.p2p_uses_nvls(Opt(True), 4) is True and dp.p2p_uses_nvls(Opt(True), 8) is True
    assert dp.p2p_uses_nvls(Opt(False), 8) is False  # no multicast address on this system


def test_shard_cells_aligned_to_chunks():
    for n, w, a in ((8000, 8, 512), (1000, 3, 128), (130, 4, 32), (5, 2, 4)):
        parts = [dp.shard_cells(n, r, w, align=a) for r in range(w)]
        cat = np.concatenate(parts)
        assert np.array_equal(cat, np.arange(n))
        assert all(p[0] % a == 0 for p in parts if p.size)
