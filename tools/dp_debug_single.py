"""One-GPU reproduction helper: train on rank `r`'s shard of a 2-rank split; variant 'full' = whole step in the graph,
'split' = data-parallel style graph (forward + backward only, optimizer eager)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cellbender_b200 import run, synth

variant = sys.argv[1]
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = "cuda:0"
ds = synth.make_config_dataset("pbmc8k", seed=0, device=dev)
a = run.default_args()
model, opt, svi, tl, te = run.setup_inference(ds, a, device=dev, rank=0, world_size=2)
model.train()
if variant == "split":
    svi.after_backward = lambda o: None
n = 0
while n < n_steps:
    for b in tl:
        svi.step(b)
        torch.cuda.synchronize()
        n += 1
        if b.n != 255 or n % 20 == 0:
            print("step", n, "B", b.n, "graphs", len(svi._graphs), flush=True)
        if n >= n_steps:
            break
print(variant, "ok", bool(torch.isfinite(opt.flat_p).all()), flush=True)
