#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/n2.err | grep "^{" > gpurun_out/n2_tailorder.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/n2_tailorder.json"))
print("N=2", round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["fwd_mpx_s"])
PY
tail -3 gpurun_out/n2.err
