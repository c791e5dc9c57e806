"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share.
  python tools/launch_summary.py gpurun_out/launches.csv [--grid] [--md]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, out = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("hp::<unnamed>::", "").replace("void ", "")
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        out.append((name, d["Grid Size"], d["Stream"], v))
    return out


def main():
    path = sys.argv[1]
    by_grid, md = "--grid" in sys.argv, "--md" in sys.argv
    launches = load(path)
    agg = collections.OrderedDict()
    tot = 0.0
    for name, grid, _, v in launches:
        key = (name, grid if by_grid and "pair_kernel" in name else "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"total {tot:.1f} us over {len(launches)} launches")
    if md:
        print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for (name, grid), (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if md:
            print(f"| {name[:60]} {grid} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
        else:
            print(f"{name[:45]:45s} {grid:16s} {n:4d} {t:9.1f} {t / n:8.1f} {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main()
