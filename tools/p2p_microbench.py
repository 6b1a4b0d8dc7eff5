"""Single-process NVLink microbenchmark of cb_clipped_adam_p2p on two GPUs (peer buffers on cuda:1, kernel on cuda:0):
DMA copy rate, remote-read-only, remote-write-only and both.  Diagnostics for dp.p2p_step."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellbender_b200 import _cabi  # noqa: E402

n = 34_703_360  # one of two shards of the PBMC8k flat buffers
d0, d1 = "cuda:0", "cuda:1"
x0 = torch.randn(n, device=d0)
x1 = torch.randn(n, device=d1)
y1 = torch.empty(n, device=d1)
y0 = torch.empty(n, device=d0)
y1.copy_(x0), y0.copy_(x1)  # enables peer access both ways
torch.cuda.synchronize(d0), torch.cuda.synchronize(d1)


def timed(fn, reps=10):
    torch.cuda.set_device(0)
    fn()
    torch.cuda.synchronize(d0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(d0)
    return e0.elapsed_time(e1) / reps


mb = 4 * n / 1e6
t = timed(lambda: y1.copy_(x0))
print(f"DMA push 0->1: {t:.4f} ms  {mb / t:.0f} GB/s")
t = timed(lambda: y0.copy_(x1))
print(f"DMA pull 1->0: {t:.4f} ms  {mb / t:.0f} GB/s")

p, m, v, g = (torch.randn(n, device=d0) for _ in range(4))
v.abs_()
lr = torch.tensor([1e-3], device=d0)
b1 = torch.tensor([0.9], device=d0)
step = torch.zeros(1, dtype=torch.int64, device=d0)
L = _cabi.lib()


def run(grads, peers):
    g_arr = (ctypes.c_void_p * len(grads))(*[t_.data_ptr() for t_ in grads])
    p_arr = (ctypes.c_void_p * max(1, len(peers)))(*[t_.data_ptr() for t_ in peers])

    def fn():
        L.call("cb_clipped_adam_p2p", p.data_ptr(), m.data_ptr(), v.data_ptr(), n, ctypes.addressof(g_arr), len(grads), 0,
               ctypes.addressof(p_arr), len(peers), lr.data_ptr(), b1.data_ptr(), 1, step.data_ptr(), 0.999, 1e-8, 10.0, 0,
               torch.cuda.current_stream().cuda_stream)
    return timed(fn)


print(f"local only             : {run([g], []):.4f} ms")
print(f"local + local 2nd grad : {run([g, x0], []):.4f} ms")
print(f"remote grad read       : {run([g, x1], []):.4f} ms")
print(f"remote param write     : {run([g], [y1]):.4f} ms")
print(f"remote read + write    : {run([g, x1], [y1]):.4f} ms")
