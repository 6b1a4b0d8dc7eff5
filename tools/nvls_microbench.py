"""torchrun microbenchmark of the data-parallel exchange kernels on symmetric buffers of the PBMC8k size (every rank runs
its 1/R shard at the same time, like a training step): cb_clipped_adam_p2p (TMA bulk peer reads + P2P stores) and
cb_clipped_adam_nvls_ex with every combination of load path (multimem.ld_reduce | local only) and store path
(multimem.st | P2P stores | local only), several grid sizes.  Milliseconds per launch, max over ranks.

    torchrun --nproc-per-node 4 tools/nvls_microbench.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from cellbender_b200 import _cabi

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
n = 69_547_264
flat_p = symm_mem.empty(n, dtype=torch.float32, device=dev).normal_()
flat_g = symm_mem.empty(n, dtype=torch.float32, device=dev).normal_()
hp = symm_mem.rendezvous(flat_p, dist.group.WORLD)
hg = symm_mem.rendezvous(flat_g, dist.group.WORLD)
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
lr, b1 = torch.tensor([1e-4], device=dev), torch.tensor([0.9], device=dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
L = _cabi.lib()
sh = n // world
a = rank * sh
f32 = torch.float32
g_ptrs = [hg.get_buffer(r, (n,), f32).data_ptr() + 4 * a for r in range(world)]
p_peers = [hp.get_buffer(r, (n,), f32).data_ptr() + 4 * a for r in range(world) if r != rank]
g_arr = (ctypes.c_void_p * world)(*g_ptrs)
p_arr = (ctypes.c_void_p * max(1, len(p_peers)))(*p_peers)
at = lambda t: t.data_ptr() + 4 * a
have_mc = bool(hg.has_multicast_support) and int(hg.multicast_ptr or 0) != 0
st = lambda: torch.cuda.current_stream().cuda_stream
tab = lambda: (lr.data_ptr(), b1.data_ptr(), 1, step.data_ptr(), 0.999, 1e-8, 10.0, 0, st())


def p2p():
    L.call("cb_clipped_adam_p2p", at(flat_p), at(m), at(v), sh, ctypes.addressof(g_arr), world, rank, ctypes.addressof(p_arr),
           len(p_peers), *tab())


def nvls(ld, stm, bps):
    def fn():
        L.call("cb_clipped_adam_nvls_ex", at(flat_p), at(m), at(v), sh, int(hg.multicast_ptr) + 4 * a, int(hp.multicast_ptr) + 4 * a,
               at(flat_g), ctypes.addressof(p_arr), len(p_peers), ld, stm, bps, *tab())
    return fn


def timed(fn, reps=10):
    for _ in range(2):
        hg.barrier(channel=0)
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    tot = 0.0
    for _ in range(reps):
        hg.barrier(channel=0)  # all ranks start together
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def say(name, ms, link_mb):
    if rank == 0:
        print(f"{name:58s} {ms:7.4f} ms   {link_mb / ms:7.0f} GB/s inbound per GPU", flush=True)


mb = 4 * sh / 1e6
if rank == 0:
    print(f"# {world} ranks, shard {sh} floats = {mb:.1f} MB, multicast {have_mc}", flush=True)
say("p2p: TMA bulk peer reads + P2P stores", timed(p2p), 2 * (world - 1) * mb)
if have_mc:
    for bps in (6, 12, 24):
        say(f"nvls: ld_reduce + multimem.st            ({bps:2d} CTAs/SM)", timed(nvls(0, 0, bps)), world * mb)
    for bps in (6, 12):
        say(f"nvls: ld_reduce + P2P stores             ({bps:2d} CTAs/SM)", timed(nvls(0, 1, bps)), world * mb)
    say("nvls: ld_reduce, local store only (reduce path alone)", timed(nvls(0, 2, 6)), mb)
    say("nvls: local grad + multimem.st (store path alone)", timed(nvls(1, 0, 6)), (world - 1) * mb)
say("local grad + P2P stores (store path alone)", timed(nvls(1, 1, 6)), (world - 1) * mb)
say("local only (HBM sweep of the shard)", timed(nvls(1, 2, 6)), 1e-9)
dist.barrier()
dist.destroy_process_group()
