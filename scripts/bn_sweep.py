"""Times the tcgen05 conv forward of the small-grid layers at several output-channel tile widths (B200LIC_TC_BN cap),
CUDA events, L2 flushed: the measurements behind the tile-width cost model in conv_tc2.cu::make_plan2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdo_ptq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
CASES = [  # name, N, Cin, H, W, Cout, k, stride, transposed
    ("h_a.4 b1", 1, 192, 16, 24, 192, 5, 2, False), ("h_a.2 b1", 1, 192, 32, 48, 192, 5, 2, False),
    ("h_a.0 b1", 1, 320, 32, 48, 192, 3, 1, False), ("g_a.6 b1", 1, 192, 64, 96, 320, 5, 2, False),
    ("g_a.4 b1", 1, 192, 128, 192, 192, 5, 2, False), ("h_s.0 b1", 1, 192, 8, 12, 320, 5, 2, True),
    ("h_s.2 b1", 1, 320, 16, 24, 480, 5, 2, True), ("h_s.4 b1", 1, 480, 32, 48, 640, 3, 1, False),
    ("g_s.0 b1", 1, 320, 32, 48, 192, 5, 2, True), ("g_s.2 b1", 1, 192, 64, 96, 192, 5, 2, True),
    ("g_a.4 b8p", 8, 192, 64, 64, 192, 5, 2, False), ("g_a.6 b8p", 8, 192, 32, 32, 320, 5, 2, False),
    ("h_a.0 b8p", 8, 320, 16, 16, 192, 3, 1, False), ("h_a.2 b8p", 8, 192, 16, 16, 192, 5, 2, False),
    ("h_s.4 b8p", 8, 480, 16, 16, 640, 3, 1, False), ("g_s.0 b8p", 8, 320, 16, 16, 192, 5, 2, True),
    ("g_a.2 b8p", 8, 192, 128, 128, 192, 5, 2, False), ("g_s.4 b8p", 8, 192, 64, 64, 192, 5, 2, True),
    ("g_s.2 b8p", 8, 192, 32, 32, 192, 5, 2, True), ("g_a.2 2K", 1, 192, 768, 1024, 192, 5, 2, False),
    ("g_s.4 2K", 1, 192, 384, 512, 192, 5, 2, True), ("g_a.4 2K", 1, 192, 384, 512, 192, 5, 2, False),
]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for name, N, Cin, H, W, Cout, k, st, tr in CASES:
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0)
    fn = (lambda: ops.deconv2d_raw(x, w, b, d)) if tr else (lambda: ops.conv2d_raw(x, w, b, d))
    row = {"case": name}
    ref = None
    VAR = "B200LIC_TC_CHAINS" if "--chains" in sys.argv else ("B200LIC_TC_MT" if "--mt" in sys.argv else "B200LIC_TC_BN")
    for cap in ((0, 1, 2, 3) if VAR.endswith("CHAINS") else ((0, 1) if VAR.endswith("MT") else (0, 256, 96, 64, 48, 32, 16))):
        os.environ[VAR] = str(cap)
        if cap == 0:
            os.environ.pop(VAR)
        for _ in range(2):
            y = fn()
        if ref is None:
            ref = y.clone()
        # GPU-bound timing: 10 calls captured in one CUDA graph (eager launches of a 3-kernel op are CPU-bound here)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                y = fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gr.replay()
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e3 / 10)
        del gr
        err = ((y - ref).norm() / ref.norm()).item()
        row["model" if cap == 0 else f"{VAR[11:].lower()}<={cap}"] = round(sorted(ts)[len(ts) // 2], 1)
        assert err < 1e-5, (name, cap, err)
    print(json.dumps(row), flush=True)
