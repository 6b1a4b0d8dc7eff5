"""torchrun check of data-parallel training at the PBMC8k shape: two epochs through the loader (full and short last
minibatch: two captured step graphs, exchange inside), progress lines per step on stderr, then the replicas are compared.

    torchrun --nproc-per-node 4 tools/dp_train_check.py [auto|p2p|sharded]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from cellbender_b200 import dp, run, synth

mode = sys.argv[1] if len(sys.argv) > 1 else None
mode = None if mode == "auto" else mode
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))


def say(*a):
    print(f"[r{rank}]", *a, file=sys.stderr, flush=True)


ds = synth.make_config_dataset("pbmc8k", seed=0, device=dev)
a = run.default_args()
model, opt, svi, tl, te = run.setup_inference(ds, a, device=dev, rank=rank, world_size=world, dp_mode=mode)
say("setup done", len(tl), "steps/epoch; n =", opt.n)
model.train()
dp.attach(svi, world, mode=mode)
say("mode", svi.dp_mode, "nvls", dp.p2p_uses_nvls(opt, world) if svi.dp_mode == "p2p" else None)
n = 0
for epoch in range(2):
    for b in tl:
        svi.step(b)
        torch.cuda.synchronize()
        n += 1
        if n <= 8 or b.n != 255:
            say("step", n, "B", b.n, "graphs", len(svi._graphs))
    say("epoch", epoch, "done after", n, "steps")
dist.barrier()
torch.cuda.synchronize()
flat = opt.flat_p.detach().clone()
others = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(others, flat)
say("replicas identical:", all(torch.equal(others[0], o) for o in others), "finite:", bool(torch.isfinite(flat).all()))
svi._graphs.clear()
torch.cuda.synchronize()
dist.destroy_process_group()
