"""A few EAGER training steps at the PBMC8k shape (no CUDA graph) for `ncu -k regex:... --set full` captures of single
kernels of the step, and (--posterior) a short posterior pass for the same purpose.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'obs_|gemm_tf32' \
        -o gpurun_out/prof python tools/step_profile.py --steps 5

The script brackets the last eager step (or, with --posterior, one pass over 2048 cells) with cudaProfilerStart / Stop.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workload", default="pbmc8k")
    ap.add_argument("--posterior", action="store_true")
    a = ap.parse_args()
    import torch

    from cellbender_b200 import run, synth

    dev = "cuda:0"
    ds = synth.make_config_dataset(a.workload, seed=0, device=dev)
    model, opt, svi, train_loader, _ = run.setup_inference(ds, run.default_args(use_cuda_graph=False), device=dev)
    model.train()
    it = iter(train_loader)
    for k in range(a.steps):
        if k == a.steps - 1 and not a.posterior:  # `ncu --profile-from-start off`: only the last eager step is captured
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        svi.step(next(it))
    torch.cuda.synchronize()
    if not a.posterior:
        torch.cuda.profiler.stop()
    if a.posterior:
        from cellbender_b200.posterior import Posterior

        model.eval()
        post = Posterior(ds, model)
        n_all = int(ds.analyzed_barcode_inds.size)
        post._latents = {"p": np.concatenate([np.ones(2048), np.zeros(n_all - 2048)])}
        post.device_posterior()  # builds the workspace, captures the chunk graphs (their warm-up runs on empty chunks)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # `ncu --profile-from-start off`: one pass over 2048 cells = four 512-cell chunks (the pipelined form)
        post.device_posterior()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
