"""-m gpu: the hand-written tcgen05 TF32 GEMM against torch (float64) on every operand-major combination and on the
shapes of the training step.  Tolerance: TF32 keeps 10 mantissa bits; the kernel is compared (a) strictly with a
float64 product of the TF32-TRUNCATED operands (same rounding as the tensor core: 2e-5 of scale, i.e. only fp32
accumulation order), (b) loosely with the exact product (2e-3 of scale)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _check(M, N, K, a_mn, b_mn, split_k=None, bias=False, accumulate=False, seed=0):
    from cellbender_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(seed)
    ld = lambda n: ops.round_up(n, 4)
    A = torch.randn((K, ld(M)) if a_mn else (M, ld(K)), device=DEV, generator=g)
    B = torch.randn((K, ld(N)) if b_mn else (N, ld(K)), device=DEV, generator=g)
    Al = (A[:, :M].t() if a_mn else A[:, :K])
    Bl = (B[:, :N].t() if b_mn else B[:, :K])
    bias_t = torch.randn(N, device=DEV, generator=g) if bias else None
    out = torch.full((M, ld(N)), 7.0, device=DEV)
    c0 = out[:, :N].clone()
    ops.gemm_tf32(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out[:, :N], bias=bias_t, accumulate=accumulate, split_k=split_k)
    torch.cuda.synchronize()
    extra = (bias_t.double() if bias else 0) + (c0.double() if accumulate else 0)
    ref_trunc = _tf32_trunc(Al.contiguous()).double() @ _tf32_trunc(Bl.contiguous()).double().t() + extra
    ref_exact = Al.double() @ Bl.double().t() + extra
    got = out[:, :N].double()
    scale = float(ref_exact.abs().max())
    e_trunc = float((got - ref_trunc).abs().max()) / scale
    e_exact = float((got - ref_exact).abs().max()) / scale
    assert float(out[:, N:].sub(7.0).abs().max() if ld(N) > N else 0.0) == 0.0, "padding columns were written"
    assert e_trunc < 2e-5, (M, N, K, a_mn, b_mn, split_k, e_trunc, e_exact)
    assert e_exact < 2e-3, (M, N, K, a_mn, b_mn, split_k, e_trunc, e_exact)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 256, 64), (200, 130, 100), (255, 512, 1000), (300, 77, 333)])
def test_gemm_all_majors_small(M, N, K, a_mn, b_mn):
    _check(M, N, K, a_mn, b_mn, split_k=1)


@pytest.mark.parametrize("split_k", [2, 7, None])
def test_gemm_split_k_bias_accumulate(split_k):
    _check(255, 512, 5000, False, False, split_k=split_k, bias=True)
    _check(255, 512, 5000, False, True, split_k=split_k, bias=True, accumulate=True)
    _check(255, 130, 999, False, False, split_k=1, bias=True, accumulate=True)


def test_gemm_cluster_multicast_variant(monkeypatch):
    """CB_GEMM_CLUSTER=1: the split-K products with four column tiles run as four-CTA clusters with the A tile TMA
    multicast; same results as the default kernel (both major orders of B, bias, accumulate, ragged K)."""
    monkeypatch.setenv("CB_GEMM_CLUSTER", "1")
    _check(255, 512, 5000, False, False, split_k=7, bias=True)
    _check(255, 512, 5000, False, True, split_k=7, bias=True, accumulate=True)
    _check(200, 400, 33538, False, False, bias=True)
    _check(255, 512, 33538, False, True)


def test_gemm_training_step_shapes():
    """PBMC8k shapes: G = 33 538 (not a multiple of 4 -> padded leading dimension), B = 255, H = 512."""
    G, Bn, H = 33538, 255, 512
    _check(Bn, H, G, False, False, bias=True)         # encoder input layers: X[B,G] . W[H,G]^T   (split-K)
    _check(Bn, G, H, False, False, split_k=1, bias=True)   # decoder output layer: h[B,H] . Wd[G,H]^T
    _check(Bn, H, G, False, True)                     # dh = dlogits[B,G] . Wd[G,H]            (B MN-major, split-K)
    _check(H, G, Bn, True, True, split_k=1)           # dW = dY^T[H,B] . X[B,G]                (both MN-major)
    _check(G, H, Bn, True, True, split_k=1)           # dWd = dlogits^T[G,B] . h[B,H]
    _check(Bn, G, H, False, True, split_k=1)          # dX = dY[B,H] . W[H,G]                  (B MN-major)


def test_gemm_two_operand_pairs_share_the_accumulator():
    """C = A1 B1^T + A2 B2^T in one launch (dx of two layers reading the same input)."""
    from cellbender_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(3)
    B, H, G = 255, 512, 33538
    dy1, dy2 = torch.randn(B, H, device=DEV, generator=g), torch.randn(B, H, device=DEV, generator=g)
    ld = ops.round_up(G + 4, 4)
    W1, W2 = torch.randn(H, ld, device=DEV, generator=g), torch.randn(H, ld, device=DEV, generator=g)
    out = ops.alloc_rows(B, G, DEV)
    ops.gemm_tf32(dy1, W1[:, 4:4 + G], B, G, H, b_mn=True, out=out[:, :G], split_k=1, a2=dy2, b2=W2[:, 4:4 + G], K2=H)
    ref = _tf32_trunc(dy1).double() @ _tf32_trunc(W1[:, 4:4 + G].contiguous()).double() \
        + _tf32_trunc(dy2).double() @ _tf32_trunc(W2[:, 4:4 + G].contiguous()).double()
    assert float((out[:, :G].double() - ref).abs().max()) / float(ref.abs().max()) < 2e-5


@pytest.mark.parametrize("M,N,off", [(512, 256, 4), (256, 512, 0), (512, 33538, 4), (33538, 512, 0), (300, 250, 0)])
def test_weight_gradient_gemm_with_fused_clipped_adam(M, N, off):
    """cb_gemm_tf32_dw_adam: W[:, off:off+N] -= Adam(dY^T X) inside the GEMM epilogue must equal the unfused sequence
    (cb_gemm_tf32 into a gradient buffer, then cb_clipped_adam_step over it), element for element; columns outside
    [off, off + N) and the padding stay untouched."""
    from cellbender_b200 import _cabi, ops

    K = 255
    g = torch.Generator(device=DEV).manual_seed(M + N)
    ld = ops.round_up(off + N, 4)
    dy = torch.randn(K, ops.round_up(M, 4), device=DEV, generator=g)
    x = torch.randn(K, ops.round_up(N, 4), device=DEV, generator=g)
    p0 = torch.randn(M, ld, device=DEV, generator=g)
    m0 = torch.randn(M, ld, device=DEV, generator=g) * 0.1
    v0 = torch.rand(M, ld, device=DEV, generator=g) * 0.01
    lr_table = torch.tensor([1e-3, 2e-3, 3e-3], device=DEV)
    b1_table = torch.tensor([0.95, 0.9, 0.85], device=DEV)
    step = torch.tensor([1], dtype=torch.int64, device=DEV)

    # unfused reference: gradient to memory, then the flat Adam sweep over the [M, N] block (row by row)
    grad = torch.zeros(M, ld, device=DEV)
    ops.gemm_tf32(dy, x, M, N, K, a_mn=True, b_mn=True, out=grad[:, off:off + N], split_k=1)
    p_ref, m_ref, v_ref = p0.clone(), m0.clone(), v0.clone()
    pr, mr, vr, gr = (t[:, off:off + N].contiguous().reshape(-1) for t in (p_ref, m_ref, v_ref, grad))
    step_ref = step.clone()
    ops.clipped_adam_step(pr, gr, mr, vr, lr_table, b1_table, step_ref, beta2=0.999, eps=1e-8, clip=10.0)
    for dst, src in ((p_ref, pr), (m_ref, mr), (v_ref, vr)):
        dst[:, off:off + N] = src.view(M, N)

    p, m, v = p0.clone(), m0.clone(), v0.clone()
    byte = 4 * off
    _cabi.lib().call("cb_gemm_tf32_dw_adam", dy.data_ptr(), dy.stride(0), 1, x.data_ptr(), x.stride(0), 1, ld, M, N, K,
                     p.data_ptr() + byte, m.data_ptr() + byte, v.data_ptr() + byte, lr_table.data_ptr(), b1_table.data_ptr(), 3,
                     step.data_ptr(), 0.999, 1e-8, 10.0, _cabi.current_stream())
    torch.cuda.synchronize()
    assert int(step) == 1  # read, not advanced
    for got, ref, name in ((p, p_ref, "param"), (m, m_ref, "exp_avg"), (v, v_ref, "exp_avg_sq")):
        diff = (got - ref).abs()
        bad = int((diff > 1e-6 * (1 + ref.abs())).sum())
        rows_bad = torch.nonzero((diff > 1e-6 * (1 + ref.abs())).any(dim=1)).flatten()[:8].tolist()
        assert bad == 0, (name, bad, float(diff.max()), rows_bad)
    assert float((p - p0).abs().max()) > 0
