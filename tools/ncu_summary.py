"""Compact per-launch table from an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`
(optionally gzipped): duration, DRAM bytes and throughput, pipe utilisation, issue slots, occupancy, registers.

    python tools/ncu_summary.py gpurun_out/s4_step_raw.csv.gz [title] > profiles/...txt
"""
import csv
import gzip
import io
import sys

COLS = [
    ("dur_us", "gpu__time_duration.sum", 1e-3, "8.1f"),
    ("dram_rd_MB", "dram__bytes_read.sum", None, "9.1f"),
    ("dram_wr_MB", "dram__bytes_write.sum", None, "9.1f"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "6.1f"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, "6.1f"),
    ("issue_%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1, "7.1f"),
    ("fma_%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1, "6.1f"),
    ("alu_%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1, "6.1f"),
    ("xu_%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1, "6.1f"),
    ("fp64_%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1, "6.1f"),
    ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, "8.1f"),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, "7.1f"),
    ("L2hit_%", "lts__t_sector_hit_rate.pct", 1, "7.1f"),
    ("regs", "launch__registers_per_thread", 1, "5.0f"),
]


def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


def to_ns(v, unit):
    u = unit.lower()
    return v * {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(u, 1.0)


def main(path, title=""):
    f = gzip.open(path) if path.endswith(".gz") else open(path, "rb")
    rows = list(csv.reader(io.TextIOWrapper(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}" if title else f"# {path}")
    print("# ncu --set full --clock-control none, one line per launch in launch order (serialised, cold-cache replays);")
    print("# *_% pipes: instructions executed on the pipe as % of its peak while the SM is active; issue_%: issue slots used;")
    print("# tensor_%: tensor-pipe cycles active; warps_%: resident warps of the maximum; dram_%: of ncu's own DRAM peak")
    print("# " + " ".join(f"{n:>{len(format(0.0, fmt))}}" for n, _, _, fmt in COLS) + "  grid        block  kernel")
    for r in data:
        out = []
        for n, key, scale, fmt in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                out.append(" " * (len(format(0.0, fmt)) - 1) + "-")
                continue
            v = float(r[ix[key]].replace(",", ""))
            u = units[ix[key]]
            if n == "dur_us":
                v = to_ns(v, u) * 1e-3
            elif n.startswith("dram_") and n.endswith("MB"):
                v = to_bytes(v, u) * 1e-6
            out.append(format(v, fmt))
        name = r[ix["Kernel Name"]].replace("void ", "").replace("cb::", "")
        print("  " + " ".join(out) + f"  {r[ix['Grid Size']].replace(' ', ''):<11} {r[ix['Block Size']].replace(' ', ''):<6} {name[:70]}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
