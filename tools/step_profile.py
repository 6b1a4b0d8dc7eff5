"""A few EAGER training steps at the PBMC8k shape (no CUDA graph) for `ncu -k regex:... --set full` captures of single
kernels of the step, and (--posterior) a short posterior pass for the same purpose.

    ncu --set full --clock-control none --import-source on -k regex:'latent_forward|elbo_obs' -s 6 -c 2 \
        -o gpurun_out/prof python tools/step_profile.py --steps 5
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
    for _ in range(a.steps):
        svi.step(next(it))
    torch.cuda.synchronize()
    if a.posterior:
        from cellbender_b200.posterior import Posterior

        model.eval()
        post = Posterior(ds, model)
        n_all = int(ds.analyzed_barcode_inds.size)
        post._latents = {"p": np.concatenate([np.ones(1024), np.zeros(n_all - 1024)])}
        post.device_posterior()
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
