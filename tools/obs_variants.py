"""A/B timing of builds of the observation kernels (cb_elbo_obs_fwd_bwd) on ONE real PBMC8k-shape minibatch: every shared
library given on the command line is loaded beside the product's and its three launches are timed back to back with CUDA
events (median of 7 measurements of 12 call sequences, L2 flushed before each measurement), and its row sums / dlogits
are compared with the product library's.

    python tools/obs_variants.py cellbender_b200/libcellbender_b200.so /path/variant1.so ...
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from cellbender_b200 import _cabi, run, synth

    dev = "cuda:0"
    ds = synth.make_config_dataset("pbmc8k", seed=0, device=dev)
    model, opt, svi, train_loader, _ = run.setup_inference(ds, run.default_args(use_cuda_graph=False), device=dev)
    model.train()
    it = iter(train_loader)
    for _ in range(3):
        svi.step(next(it))
    batch = next(it)
    inp = batch.encoder_inputs(svi._ambient_unit_vector())
    with torch.no_grad():
        model.elbo_terms(batch.loader.csr, batch.row_ids, inp)
    csr_, rows_, bufs_ = model._obs_ctx
    B, G = batch.size(0), model.n_genes
    torch.manual_seed(0)
    coef = torch.rand((B, 7), device=dev) * torch.tensor([3000, 100, 30, 100, 3, 100, 3], device=dev)
    wts, al = -torch.rand((B, 3), device=dev), torch.tensor([5.0], device=dev)
    oo = bufs_["obs_out"]
    ptr = _cabi.ptr
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = _cabi.current_stream()
    ref = None
    for path in sys.argv[1:]:
        dll = ctypes.CDLL(os.path.abspath(path))
        dll.cb_elbo_obs_scratch_elems.restype = ctypes.c_longlong
        scratch = torch.empty(int(dll.cb_elbo_obs_scratch_elems(ctypes.c_int(B))), device=dev)
        sums, dlog, qamb = torch.zeros_like(oo["sums"]), torch.zeros_like(oo["dlogits"]), torch.zeros_like(oo["qamb"])
        V, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        fn = dll.cb_elbo_obs_fwd_bwd
        fn.restype = ctypes.c_int
        fn.argtypes = [V, V, V, V, I, I, V, L, V, V, V, V, V, V, V, V, L, V, V, V, V]
        ld = bufs_["logits"].stride(0)
        args = (ptr(csr_.indptr), ptr(csr_.indices), ptr(csr_.data), ptr(rows_), B, G, ptr(bufs_["logits"]), ld, ptr(bufs_["chi_a"]),
                ptr(bufs_["chi_b"]), ptr(coef), ptr(al), ptr(wts), ptr(sums), ptr(dlog), ptr(qamb), ld, ptr(oo["lse"]),
                ptr(oo["nan_flag"]), ptr(scratch), st)
        assert fn(*args) == 0
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0.record()
            for _ in range(12):
                fn(*args)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 12 * 1e3)
        got = (sums.double().cpu().numpy(), dlog[:, :G].double().cpu().numpy(), qamb[:, :G].double().cpu().numpy())
        if ref is None:
            ref = got
        err = [float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) for a, b in zip(got, ref)]
        print(f"{os.path.basename(path):40s} {np.median(ts):7.1f} us per call (min {min(ts):.1f})   max |diff| / max |ref| of "
              f"sums {err[0]:.1e}, dlogits {err[1]:.1e}, qamb {err[2]:.1e}", flush=True)
    model.nan_flag.zero_()


if __name__ == "__main__":
    main()
