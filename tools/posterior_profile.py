"""Run the posterior pass (encoder once per chunk, S latent samples, decoder GEMM, noise-count window, COO) on the first
``--cells`` droplets of the PBMC8k-shape synthetic dataset with a randomly initialised model.  Used under
`ncu --metrics gpu__time_duration.sum` for the launch list of the posterior phase (profiles/), and stand-alone for
CUDA-event timings of the phase.

    python tools/posterior_profile.py --cells 2048 --reps 3
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--workload", default="pbmc8k")
    ap.add_argument("--estimators", action="store_true")
    ap.add_argument("--timeline", default=None, help="write a per-kernel timeline of one pass (torch.profiler / CUPTI)")
    ap.add_argument("--sweep", action="store_true", help="time the pipelined pass for several (tile_cfg, window CTAs per SM)")
    a = ap.parse_args()
    import torch

    from cellbender_b200 import run, synth
    from cellbender_b200.posterior import Posterior

    dev = "cuda:0"
    torch.manual_seed(0)  # the model is randomly initialised: same weights (hence the same COO size) in every run
    ds = synth.make_config_dataset(a.workload, seed=0, device=dev)
    model = run.build_model(ds, run.default_args(), device=dev)
    model.eval()
    model.encoder["other"].offset = {"logit_p": 0.0, "epsilon": 0.0}
    post = Posterior(ds, model, posterior_batch_size=128)
    n_all = int(ds.analyzed_barcode_inds.size)
    n_cells = min(a.cells, n_all)
    post._latents = {"p": np.concatenate([np.ones(n_cells), np.zeros(n_all - n_cells)]), "z": None, "d": None,
                     "epsilon": None, "phi_loc_scale": None}
    post.device_posterior(cell_subset=np.arange(min(128, n_cells)))
    torch.cuda.synchronize()
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dp_ = post.device_posterior()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"posterior: {n_cells} cells, {int(dp_.m.numel())} COO entries, {dt * 1e3:.2f} ms, {n_cells / dt:.0f} droplets/s",
              flush=True)
    if a.sweep:
        for cfg, ctas in ((0, 0), (3, 0), (3, 1), (3, 2), (3, 3), (0, 2), (1, 2), (3, 4)):
            post.pipeline_tile_cfg, post.pipeline_window_ctas = cfg, ctas
            post._chunk_graph = None
            post.device_posterior()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                post.device_posterior()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            print(f"tile_cfg {cfg}, window CTAs/SM {ctas}: {1e3 * min(ts):.2f} ms, {n_cells / min(ts):.0f} droplets/s", flush=True)
    if a.timeline:
        import collections
        import json
        import tempfile

        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            post.device_posterior()
            torch.cuda.synchronize()
        path = os.path.join(tempfile.mkdtemp(), "trace.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
        ev.sort(key=lambda e: e["ts"])
        t0, t1 = ev[0]["ts"], ev[-1]["ts"] + ev[-1]["dur"]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for e in ev:
            n = e["name"].replace("void ", "").split("(")[0][:100]
            agg[n][0] += 1
            agg[n][1] += e["dur"]
        busy = sum(v[1] for v in agg.values())
        lines = [f"# posterior pass, {n_cells} cells: {len(ev)} kernels, span {t1 - t0:.0f} us, sum of kernel durations {busy:.0f} us "
                 f"({100 * busy / (t1 - t0):.0f} % of the span: the rest is launch gaps / host)", "# total_us  launches  kernel"]
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            lines.append(f"{t:10.1f} {c:6d}  {n}")
        # raw per-stream timeline of the middle of the pass (two pipeline iterations): what overlaps what
        big = [e for e in ev if "posterior_window" in e["name"]]
        if len(big) >= 8:
            w0, w1 = big[len(big) // 2]["ts"], big[len(big) // 2 + 2]["ts"]
            streams = sorted({e["args"].get("stream", 0) for e in ev})
            lines.append("# ---- two iterations from the middle of the pass: start_us  dur_us  stream  kernel (>= 8 us only)")
            for e in ev:
                if w0 <= e["ts"] < w1 and e["dur"] >= 8:
                    lines.append(f"{e['ts'] - w0:9.1f} {e['dur']:8.1f}  s{streams.index(e['args'].get('stream', 0))}  "
                                 f"{e['name'].replace('void ', '').split('(')[0][:80]}")
        open(a.timeline, "w").write("\n".join(lines) + "\n")
        print("\n".join(lines[:12]))
    if a.estimators:
        from cellbender_b200.estimation import Mean, MultipleChoiceKnapsack

        conv = post.index_converter
        for _ in range(2):
            t0 = time.perf_counter()
            mean_csr = Mean(index_converter=conv).estimate_noise(dp_, None, device=dev)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            tgt = np.asarray(mean_csr.sum(axis=0)).ravel().astype(np.float64)
            noise = MultipleChoiceKnapsack(index_converter=conv).estimate_noise(dp_, None, noise_targets_per_gene=tgt, device=dev)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"Mean {1e3 * (t1 - t0):.1f} ms, MCKP {1e3 * (t2 - t1):.1f} ms, removed {int(noise.sum())}", flush=True)


if __name__ == "__main__":
    main()
