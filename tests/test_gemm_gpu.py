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


def _check(M, N, K, a_mn, b_mn, split_k=None, bias=False, accumulate=False, seed=0, tile_cfg=0):
    from cellbender_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(seed)
    ld = lambda n: ops.round_up(n, 4)
    A = torch.randn((K, ld(M)) if a_mn else (M, ld(K)), device=DEV, generator=g)
    B = torch.randn((K, ld(N)) if b_mn else (N, ld(K)), device=DEV, generator=g)
    Al = (A[:, :M].t() if a_mn else A[:, :K])
    Bl = (B[:, :N].t() if b_mn else B[:, :K])
    bias_t = torch.randn(N, device=DEV, generator=g) if bias else None
    out = torch.full((M + 2, ld(N)), 7.0, device=DEV)  # two guard rows behind the matrix
    c0 = out[:M, :N].clone()
    ops.gemm_tf32(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out[:M, :N], bias=bias_t, accumulate=accumulate, split_k=split_k,
                  tile_cfg=tile_cfg)
    torch.cuda.synchronize()
    extra = (bias_t.double() if bias else 0) + (c0.double() if accumulate else 0)
    ref_trunc = _tf32_trunc(Al.contiguous()).double() @ _tf32_trunc(Bl.contiguous()).double().t() + extra
    ref_exact = Al.double() @ Bl.double().t() + extra
    got = out[:M, :N].double()
    scale = float(ref_exact.abs().max())
    e_trunc = float((got - ref_trunc).abs().max()) / scale
    e_exact = float((got - ref_exact).abs().max()) / scale
    # TMA stores clip rows exactly and columns at 16-byte granularity: the padding columns of the last 16-byte chunk of a
    # row may receive zeros (products with zero-filled operand rows), nothing else is touched
    assert float(out[M:].sub(7.0).abs().max()) == 0.0, "rows behind the matrix were written"
    if ld(N) > N:
        pad = out[:M, N:]
        assert bool(((pad == 7.0) | (pad == 0.0)).all()), "padding columns hold data"
        assert float(out[:M, (N + 3) // 4 * 4:].sub(7.0).abs().max() if ld(N) > (N + 3) // 4 * 4 else 0.0) == 0.0
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


@pytest.mark.parametrize("tile_cfg", [1, 2, 3, 4, 5, 6, 7])
def test_gemm_every_tile_configuration(tile_cfg):
    """cb_gemm_tf32_ex: every tile configuration (256 x 128 / 256 x 256 / 128 x 256 / 128 x 128, 2-stage double-CTA and deep
    rings) gives the same product, with bias, accumulation (TMA reduce-add), split-K and ragged edges (TMA-store clipping)."""
    _check(255, 512, 1000, False, False, split_k=1, bias=True, tile_cfg=tile_cfg)
    _check(255, 1030, 255, True, True, split_k=1, tile_cfg=tile_cfg)
    _check(200, 333, 999, False, True, split_k=3, bias=True, accumulate=True, tile_cfg=tile_cfg)
    _check(300, 77, 333, True, False, split_k=1, accumulate=True, tile_cfg=tile_cfg)


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


def test_gemm_rejects_operands_tma_cannot_address():
    """No library fallback: an output / operand with a row pitch that is not a multiple of 16 bytes is an error."""
    from cellbender_b200 import _cabi, gemm, ops

    a = torch.randn(64, 33, device=DEV)
    b = torch.randn(64, 33, device=DEV)
    with pytest.raises((AssertionError, _cabi.CabiError)):
        ops.gemm_tf32(a, b, 64, 64, 33)
    lin = torch.nn.Linear(33, 64).to(DEV)
    with pytest.raises(RuntimeError, match="TMA"):
        gemm.linear_big(a, lin, n_in=33)  # the weight's 132-byte rows: not addressable until re-homed
    gemm.pad_big_weights(lin)
    assert lin.weight.stride(0) == 36 and lin.weight.shape == (64, 33)
    y = gemm.linear_big(a, lin, n_in=33)  # the small activation matrix is zero-padded on the fly
    ref = torch.nn.functional.linear(a.double(), lin.weight.double(), lin.bias.double())
    assert float((y.double() - ref).abs().max()) < 2e-2
    big = torch.nn.Linear(33538 + 4, 512).to(DEV)
    gemm.pad_big_weights(big)
    assert big.weight.stride(0) % 32 == 0 and big.weight.shape == (512, 33542)


def test_feat_dw_matches_matmul():
    from cellbender_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(11)
    dy = torch.randn(255, 512, device=DEV, generator=g)
    feat = torch.randn(255, 4, device=DEV, generator=g)
    dw = torch.full((512, 36), 7.0, device=DEV)
    ops.feat_dw(dy, feat, dw)
    ref = (dy.double().t() @ feat.double()).float()
    torch.testing.assert_close(dw[:, :4], ref, rtol=1e-5, atol=1e-4)
    assert float((dw[:, 4:] - 7.0).abs().max()) == 0.0


def test_captured_split_k_product_survives_later_larger_workspaces():
    """A captured CUDA graph refers to the split-K workspace by address.  A later product that needs a LARGER workspace
    (the smaller last minibatch of an epoch picks more splits) followed by torch.cuda.empty_cache() -- which every
    torch.cuda.graph() capture performs -- must leave the first graph replayable: the workspace may not be a cached buffer
    that gets replaced and freed (found as an illegal address at the first replay after the second capture)."""
    from cellbender_b200 import ops

    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(3)
    a = torch.randn((64, 4096), device=dev, generator=g)
    b = torch.randn((128, 4096), device=dev, generator=g)
    out = torch.zeros((64, 128), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.gemm_tf32(a, b, 64, 128, 4096, out=out, split_k=4)
    torch.cuda.current_stream().wait_stream(side)
    want = out.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.gemm_tf32(a, b, 64, 128, 4096, out=out, split_k=4)
    out.zero_()
    graph.replay()
    assert torch.equal(out, want)
    # a product with a much larger workspace, then a second capture (which empties the allocator's cache)
    a2 = torch.randn((256, 8192), device=dev, generator=g)
    b2 = torch.randn((512, 8192), device=dev, generator=g)
    ops.gemm_tf32(a2, b2, 256, 512, 8192, split_k=16)
    graph2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph2):
        ops.gemm_tf32(a2, b2, 256, 512, 8192, split_k=16)
    torch.cuda.empty_cache()
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
