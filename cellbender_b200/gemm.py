"""The four large contractions of a training step (three G -> 512 encoder input layers, one 512 -> G decoder
output layer; reference vae/encoder.py:97-105,211,220 and vae/decoder.py:41-47).

``BACKEND``:
  'cublas'   torch.matmul (library GEMM; fp32 unless TF32 is allowed) -- the plumbing default while the
             hand-written kernel is being brought up;
  'tcgen05'  csrc/gemm_tf32.cu: TMA-fed tcgen05.mma kind::tf32 with TMEM accumulators (set by
             ``cellbender_b200.gemm.set_backend('tcgen05')`` once its parity tests are green on the GPU).
"""
from typing import Optional

import torch
import torch.nn.functional as F

BACKEND = "cublas"


def set_backend(name: str):
    global BACKEND
    assert name in ("cublas", "tcgen05")
    BACKEND = name


def linear_big(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], n_in: int, col_offset: int = 0,
               wide_output: bool = False) -> torch.Tensor:
    """y = x[:, :n_in] @ weight[:, col_offset:col_offset + n_in].T + bias   (nn.Linear layout: weight [out, in]).

    x may carry padding columns beyond n_in (rows are 16-byte aligned buffers from the gather kernel)."""
    xs = x[:, :n_in]
    w = weight[:, col_offset:col_offset + n_in]
    return F.linear(xs, w, bias)
