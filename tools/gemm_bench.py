"""Times the large contractions of one PBMC8k-shape training step (and the posterior's decoder product) in isolation, for
every tile configuration of cb_gemm_tf32_ex, with an L2 flush between launches (CUDA events on the launching stream).
Reports algorithmic HBM bytes / time against MEASURED_PEAKS.json.  Used to choose the heuristic in csrc/gemm_tf32.cu.

    python tools/gemm_bench.py > gpurun_out/gemm_bench.txt
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cellbender_b200 import ops  # noqa: E402

DEV = "cuda:0"
G, B, H, Z = 33538, 255, 512, 64
ldG = ops.round_up(G + 4, 4)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0}
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
CFG = {0: "auto", 1: "256x128 s2x2", 2: "256x128 deep", 3: "256x256 deep", 4: "128x256 s2x2", 5: "128x256 deep", 6: "128x128 s2x2",
       7: "128x128 deep"}


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def case(name, make, alg_bytes, cfgs, splits=(None,)):
    print(f"== {name}: algorithmic {alg_bytes / 1e6:.1f} MB -> {alg_bytes / peaks['hbm_gbs'] / 1e3:.1f} us at {peaks['hbm_gbs']:.0f} GB/s")
    for cfg in cfgs:
        for sk in splits:
            try:
                us = timed(lambda: make(cfg, sk))
                print(f"   cfg {cfg} ({CFG[cfg]:13s}) split_k {str(sk):5s}: {us:7.1f} us   {alg_bytes / us / 1e3:7.0f} GB/s   "
                      f"frac {alg_bytes / us / 1e3 / peaks['hbm_gbs']:.2f}")
            except Exception as e:  # noqa: BLE001
                print(f"   cfg {cfg} split_k {sk}: FAILED {type(e).__name__}: {str(e)[:100]}")


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    x = rn(B, ldG)            # activations [B, G] (padded rows)
    W = rn(H, ldG)            # encoder weight [H, G + 4]
    W2 = rn(H, ldG)
    Wd = rn(G, H)             # decoder weight [G, H]
    h = rn(B, H)
    dy, dy2 = rn(B, H), rn(B, H)
    dl = rn(B, ldG)           # dlogits [B, G]
    bias = rn(G)
    out_bh = ops.alloc_rows(B, H, DEV)
    out_bg = ops.alloc_rows(B, G, DEV)
    dW = torch.zeros(H, ldG, device=DEV)
    dWd = torch.zeros(G, H, device=DEV)

    case("encoder forward  X[B,G] . W[H,G]^T (split-K)", lambda c, s: ops.gemm_tf32(x, W[:, 4:4 + G], B, H, G, out=out_bh[:, :H], split_k=s, tile_cfg=c),
         4.0 * G * H + 4.0 * B * G, [0, 2, 3, 5], splits=(None, 37, 74))
    case("decoder forward  h[B,H] . Wd[G,H]^T + bias", lambda c, s: ops.gemm_tf32(h, Wd, B, G, H, out=out_bg[:, :G], bias=bias, split_k=1, tile_cfg=c),
         4.0 * G * H + 4.0 * B * G, [0, 1, 2, 3])
    case("dh = dlogits[B,G] . Wd[G,H] (split-K)", lambda c, s: ops.gemm_tf32(dl, Wd, B, H, G, b_mn=True, out=out_bh[:, :H], split_k=s, tile_cfg=c),
         4.0 * G * H + 4.0 * B * G, [0, 2, 3], splits=(None, 37, 74))
    case("dW = dY^T[H,B] . X[B,G]", lambda c, s: ops.gemm_tf32(dy, x, H, G, B, a_mn=True, b_mn=True, out=dW[:, 4:4 + G], split_k=1, tile_cfg=c),
         4.0 * G * H + 4.0 * B * G, [0, 1, 2, 3, 4, 5])
    case("dWd = dlogits^T[G,B] . h[B,H]", lambda c, s: ops.gemm_tf32(dl, h, G, H, B, a_mn=True, b_mn=True, out=dWd, split_k=1, tile_cfg=c),
         4.0 * G * H + 4.0 * B * G, [0, 1, 2, 3, 4, 5])
    case("dX pair = dY1 . W1 + dY2 . W2", lambda c, s: ops.gemm_tf32(dy, W[:, 4:4 + G], B, G, H, b_mn=True, out=out_bg[:, :G], split_k=1,
                                                                     a2=dy2, b2=W2[:, 4:4 + G], K2=H, tile_cfg=c),
         8.0 * G * H + 4.0 * B * G, [0, 1, 2, 3])
    # small layers (latency-bound): [B, 512] x [512, 512], [B, 64] x [512, 64]
    Ws, hs = rn(H, H), rn(B, H)
    case("hidden layer  h[B,512] . W[512,512]^T", lambda c, s: ops.gemm_tf32(hs, Ws, B, H, H, out=out_bh[:, :H], split_k=s, tile_cfg=c),
         4.0 * H * H + 8.0 * B * H, [0, 1, 2, 5, 7], splits=(1, 2, 4))
    Wz, z = rn(H, Z), rn(B, Z)
    case("decoder hidden  z[B,64] . W[512,64]^T", lambda c, s: ops.gemm_tf32(z, Wz, B, H, Z, out=out_bh[:, :H], split_k=1, tile_cfg=c),
         4.0 * H * Z + 4.0 * B * (H + Z), [0, 1, 2, 7])
    # posterior decoder product, transposed output: W[G,512] . h[S*Bc,512]^T -> [G, 10240]
    n = 20 * 512
    hp = rn(n, H)
    lt = torch.empty(G, n, device=DEV)
    flops = 2.0 * G * n * H
    for cfg in (0, 1, 2, 3):
        us = timed(lambda: ops.gemm_tf32(Wd, hp, G, n, H, out=lt, split_k=1, tile_cfg=cfg), reps=3)
        print(f"== posterior decoder [G,512] x [10240,512]^T cfg {cfg} ({CFG[cfg]}): {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  "
              f"(write {4.0 * G * n / us / 1e3:.0f} GB/s)")
    # Adam sweep (grid capped at 4 CTAs per SM)
    P = 69_406_720
    p_, g_, m_, v_ = (torch.randn(P, device=DEV) for _ in range(4))
    v_.abs_()
    lr = torch.tensor([1e-4], device=DEV)
    b1 = torch.tensor([0.9], device=DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    us = timed(lambda: ops.clipped_adam_step(p_, g_, m_, v_, lr, b1, step), reps=5)
    print(f"== clipped_adam_step {P} params: {us:.1f} us  {28.0 * P / us / 1e3:.0f} GB/s  frac {28.0 * P / us / 1e3 / peaks['hbm_gbs']:.2f}")


if __name__ == "__main__":
    main()
