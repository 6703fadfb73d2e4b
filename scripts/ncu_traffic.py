"""profiles/traffic.json from an `ncu --set full` raw CSV export: per kernel (first launch matching each key) the DRAM
bytes read + written, the duration and the tensor-pipe activity.  bench.py reads `traffic` for its roofline from here.

  ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > profiles/r2_ncu_full_raw.csv
  python scripts/ncu_traffic.py profiles/r2_ncu_full_raw.csv "tc2_gather_gemm_kernel g_a.2=tc2_gather_gemm_kernel:0" ...
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "ns",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__grid_size": "grid",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    H, U = rows[hdr], rows[hdr + 1]
    ki = H.index("Kernel Name")
    cols = {H.index(k): (v, U[H.index(k)]) for k, v in WANT.items() if k in H}
    launches = []
    for r in rows[hdr + 2:]:
        if len(r) <= ki:
            continue
        e = {"kernel": r[ki].split("(")[0].replace("b200lic::", "").replace("void ", "")}
        for i, (name, unit) in cols.items():
            try:
                e[name] = float(r[i].replace(",", "")) * UNIT.get(unit, 1.0)
            except ValueError:
                pass
        launches.append(e)
    out = {}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        out = json.load(open(path))
    for spec in sys.argv[2:]:
        key, sel = spec.split("=")
        name, idx = sel.split(":")
        hits = [l for l in launches if l["kernel"].startswith(name)]
        if len(hits) > int(idx):
            l = hits[int(idx)]
            out[key] = l.get("dram_read", 0.0) + l.get("dram_write", 0.0)
            out[key + " detail"] = l
    json.dump(out, open(path, "w"), indent=1)
    for l in launches:
        print(json.dumps(l))


if __name__ == "__main__":
    main()
