"""cProfile of the estimation + output-assembly phase (run.compute_output_denoised_counts) and of latents_map at the
PBMC8k shape with a randomly initialised model: where the host time of the post-training phases goes."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cellbender_b200 import run, synth
from cellbender_b200.posterior import Posterior

dev = "cuda:0"
torch.manual_seed(0)
ds = synth.make_config_dataset("pbmc8k", seed=0, device=dev)
args = run.default_args()
model = run.build_model(ds, args, device=dev)
model.eval()
model.encoder["other"].offset = {"logit_p": 0.0, "epsilon": 0.0}
post = Posterior(ds, model, posterior_batch_size=128)
t0 = time.perf_counter()
_ = post.latents_map
torch.cuda.synchronize()
print(f"latents_map {time.perf_counter() - t0:.3f} s")
n_all = int(ds.analyzed_barcode_inds.size)
post._latents["p"] = np.concatenate([np.ones(8000), np.zeros(n_all - 8000)])
t0 = time.perf_counter()
post.cell_noise_count_posterior_coo()
print(f"posterior coo (scipy on host) {time.perf_counter() - t0:.3f} s")
run.compute_output_denoised_counts(post, args)  # warm-up
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
run.compute_output_denoised_counts(post, args)
torch.cuda.synchronize()
pr.disable()
print(f"compute_output_denoised_counts {time.perf_counter() - t0:.3f} s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
