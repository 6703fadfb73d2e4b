"""Fused (cluster) vs three-launch dynamic activation quantiser, graph-replay timing with L2 flushed."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdo_ptq_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)
for shape in [(1, 192, 32, 48), (1, 192, 64, 96), (1, 192, 128, 192), (1, 192, 256, 384), (1, 192, 768, 1024), (1, 320, 96, 128),
              (8, 192, 64, 64), (1, 96, 128, 192)]:
    x = torch.randn(shape, device=dev)
    row = {"shape": shape}
    for fused in (True, False):
        ops.ACTQ_FUSED = "always" if fused else False
        ops.act_quant(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                y = ops.act_quant(x)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 100)
        row["fused_us" if fused else "three_us"] = round(sorted(ts)[2], 1)
    row["GBps_fused"] = round(x.numel() * 8 / row["fused_us"] / 1e3, 1)
    print(json.dumps(row))
ops.ACTQ_FUSED = True
