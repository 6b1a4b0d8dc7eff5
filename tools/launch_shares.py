"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel shares of one training step
(delimited by two consecutive gather_rows_kernel launches: the first kernel of every step)."""
import collections
import csv
import re
import sys


def main(path, out=None, title=""):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "gather_rows_kernel" in n]

    def val(r):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        return v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)

    # the last complete step: the longest run between two consecutive gathers among the last few (the bench's e2e loop
    # and tail launch other gathers beside the step's own)
    a, b = max(list(zip(idx[:-1], idx[1:]))[-12:], key=lambda ab: ab[1] - ab[0])
    seg = rows[a:b]
    tot = sum(val(r) for r in seg)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"void ", "", n)[:110]
        agg[n][0] += 1
        agg[n][1] += val(r)
    lines = [f"# {title}", f"# kernels in the step: {len(seg)}   sum of kernel durations: {tot:.1f} us "
             "(ncu: cold-cache, serialised -> compare shares, not absolutes)", "# share_%   total_us   launches   kernel"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"{100 * t / tot:7.2f}  {t:9.1f}  {c:5d}   {n}")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "")
