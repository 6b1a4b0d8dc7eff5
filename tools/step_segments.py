"""Print the kernels of one training step (between two clipped_adam launches of consecutive steps) from an
`ncu --metrics gpu__time_duration.sum --csv` launch list, torch kernels grouped between ours."""
import csv
import re
import sys


def main(path, verbose=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    # one advance_step_kernel per optimizer step closes a step; the roofline loop of bench.py at the end launches it
    # back to back, so take the last pair that encloses a whole step
    idx = [i for i, n in enumerate(names) if "advance_step" in n]
    gaps = [(idx[i], idx[i + 1] - idx[i]) for i in range(len(idx) - 1)]
    big = [g for g in gaps if g[1] > 50]
    a = big[-1][0] + 1
    b = a + big[-1][1]
    seg = rows[a:b]

    def val(r):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        return v / 1e3 if u.startswith("n") else v

    cnt = t = 0
    tot = ours = 0.0
    for r in seg:
        n = r["Kernel Name"]
        tot += val(r)
        if "cb::" in n[:14]:
            ours += val(r)
            if cnt:
                print(f"   ... {cnt} torch kernels, {t:.0f} us")
            print(f"{val(r):8.1f} us  {r['Grid Size']:>14}  {re.sub(r'[(].*', '', n)[:60]}")
            cnt = t = 0
        else:
            cnt += 1
            t += val(r)
            if verbose:
                short = re.sub(r"void at::native::|at::native::", "", n)
                print(f"      {val(r):7.1f} {r['Grid Size']:>14} {short[:100]}")
    if cnt:
        print(f"   ... {cnt} torch kernels, {t:.0f} us")
    print(f"# kernels {len(seg)}  sum {tot:.1f} us  ours {ours:.1f} us  torch {tot - ours:.1f} us")
    import collections

    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        n = re.sub(r"[(].*", "", r["Kernel Name"]).replace("void ", "")
        n = re.sub(r"at::native::|<unnamed>::", "", n)[:70]
        agg[n][0] += 1
        agg[n][1] += val(r)
    print("# share_%   total_us   launches   kernel")
    for n, (c, t_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{100 * t_ / tot:7.2f}  {t_:9.1f}  {c:5d}   {n}")


if __name__ == "__main__":
    main(sys.argv[1], verbose=len(sys.argv) > 2)
