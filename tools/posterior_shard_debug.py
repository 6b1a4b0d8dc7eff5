"""torchrun helper: per-rank timings of barcode-sharded posterior passes (every pass printed)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from cellbender_b200 import dp, run, synth
from cellbender_b200.posterior import Posterior

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
ds = synth.make_config_dataset("pbmc8k", seed=0, device=dev)
model = run.build_model(ds, run.default_args(), device=dev)
model.eval()
model.encoder["other"].offset = {"logit_p": 0.0, "epsilon": 0.0}
post = Posterior(ds, model, posterior_batch_size=128)
n_all, n_cells = int(ds.analyzed_barcode_inds.size), 8000
post._latents = {"p": np.concatenate([np.ones(n_cells), np.zeros(n_all - n_cells)])}
shard = dp.shard_cells(n_cells, rank, world, align=post.cells_per_chunk)
post.device_posterior(cell_subset=np.arange(128))
for i in range(6):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    piece = post.device_posterior(cell_subset=shard)
    torch.cuda.synchronize()
    print(f"[r{rank}] pass {i}: {shard.size} cells, {piece.m.numel()} entries, {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
dist.barrier()
dist.destroy_process_group()
