#!/usr/bin/env bash
# Back-to-back bench of library variants on ONE box (box-to-box spread is ~3 %, larger than most kernel changes):
#   gpurun -- 'bash scripts/ab_bench.sh A old A old'      # A = the in-tree library, other names = build/variants/<name>
# Prints ms per calibration sweep, streaming e2e imgs/s and the isolated conv-forward launch time per variant.
set -u
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/gpurun_out"
extra=${AB_BENCH_ARGS:---skip-cpu --skip-fwd --steps 20 --warmup 3}
for v in "$@"; do
  if [ "$v" = A ]; then unset B200LIC_LIB; else export B200LIC_LIB="$root/build/variants/$v/libb200lic.so"; fi
  timeout 300 python "$root/bench.py" $extra 2> "$root/gpurun_out/ab_$v.err" > "$root/gpurun_out/ab_$v.json"
  python - "$v" "$root/gpurun_out/ab_$v.json" <<'PY'
import json, sys
v, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(v, "ms/step", round(d["ms_per_step"], 4), "calib imgs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "conv launch ms", round(d["roofline"]["ms_per_launch"], 4), "fwd", d.get("fwd_mpx_s"), d.get("fwd_mpx_s_2k"))
except Exception as e:
    print(v, "FAILED:", e)
PY
done
