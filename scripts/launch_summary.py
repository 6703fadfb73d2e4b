"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time per kernel and, for a calibration sweep
captured by scripts/sweep_launches.py, per reconstruction unit (units start at a gather_mix / stage_mix launch).

  python scripts/launch_summary.py gpurun_out/launches_sweep.csv [--order]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    H = rows[hdr]
    ki, mi, vi, gi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Grid Size")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) > vi and r[mi] == "gpu__time_duration.sum":
            name = r[ki].split("(")[0].replace("b200lic::", "").replace("void ", "")
            out.append((name, float(r[vi].replace(",", "")) / 1e3, r[gi]))
    return out


def main():
    data = load(sys.argv[1])
    tot = sum(d[1] for d in data)
    print(f"{len(data)} launches, {tot:.1f} us serialised")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, _ in data:
        agg[k][0] += 1
        agg[k][1] += v
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k[:56]:56s} {n:4d} {v:9.1f} us {100 * v / tot:5.1f} %")
    if "--order" in sys.argv:
        unit, acc = -1, 0.0
        for k, v, g in data:
            if k.startswith("gather_mix") or k.startswith("stage_mix"):
                if unit >= 0:
                    print(f"  -- unit {unit}: {acc:.1f} us")
                unit += 1
                acc = 0.0
            acc += v
            print(f"{k[:48]:48s} {v:8.1f} {g}")
        print(f"  -- unit {unit}: {acc:.1f} us")


if __name__ == "__main__":
    main()
