"""Per-stream kernel timeline of ONE CUDA-graph-replayed training step (PBMC8k shape) from torch.profiler (CUPTI activity
records: start time, duration, stream of every kernel inside the replayed graph).  ncu serialises kernels, so it cannot
show what overlaps what; this does.  Output: a text table sorted by start time, one line per kernel.

    python tools/step_timeline.py --out gpurun_out/step_timeline.txt
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pbmc8k")
    ap.add_argument("--out", default="gpurun_out/step_timeline.txt")
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    from cellbender_b200 import run, synth

    # under torchrun: the data-parallel step (rank 0 writes its timeline)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist

        from cellbender_b200 import dp

        dist.init_process_group("nccl", device_id=torch.device(dev))
    ds = synth.make_config_dataset(a.workload, seed=0, device=dev)
    model, opt, svi, train_loader, _ = run.setup_inference(ds, run.default_args(), device=dev, rank=rank, world_size=world)
    model.train()
    if world > 1:
        dp.attach(svi, world)
    batches = []
    while len(batches) < 12 + a.steps:
        try:
            batches.append(next(train_loader))
        except StopIteration:
            continue
    for b in batches[:12]:
        svi.step(b)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for b in batches[12:]:
            svi.step(b)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        svi._graphs.clear()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        if rank != 0:
            return
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ks.sort(key=lambda e: e["ts"])
    # split into steps at the first kernel of each replay: gather_rows_kernel launches once per step
    starts = [i for i, e in enumerate(ks) if "gather_rows" in e["name"]]
    lines = []
    if len(starts) >= 3:
        # a complete step: the longest run of kernels between two consecutive gathers (CUPTI can drop the tail of the
        # last replay; the first replay inside the profiler carries its start-up cost)
        bounds = list(zip(starts[:-1], starts[1:]))
        i0, i1 = max(bounds[1:] or bounds, key=lambda b: b[1] - b[0])
        seg = ks[i0:i1]
        starts = [i0, i1]
    else:
        seg = ks
    t0 = seg[0]["ts"]
    streams = sorted({e["args"].get("stream", 0) for e in seg})
    lines.append(f"# one replayed step: {len(seg)} kernels on {len(streams)} streams, span "
                 f"{seg[-1]['ts'] + seg[-1]['dur'] - t0:.1f} us (next step's first kernel starts at "
                 f"{ks[starts[-1]]['ts'] - t0:.1f} us)")
    lines.append("# start_us   dur_us  stream  kernel")
    for e in seg:
        name = e["name"].replace("void ", "")
        name = name.split("(")[0][:90]
        lines.append(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f}  s{streams.index(e['args'].get('stream', 0))}  {name}")
    busy = {}
    for e in seg:
        s = streams.index(e["args"].get("stream", 0))
        busy[s] = busy.get(s, 0.0) + e["dur"]
    lines.append("# busy us per stream: " + ", ".join(f"s{k}: {v:.1f}" for k, v in sorted(busy.items())))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:3]))
    print(f"wrote {a.out}")


if __name__ == "__main__":
    main()
