"""A/B of the CTA-pair form of the conv engine (b200lic_set_option("pair", 0 / 2)): correctness of both forms against an
fp64 convolution and time per op (10 calls in one CUDA graph, L2 flushed between replays).  Each line is flushed as soon
as it is known so that a hang shows where it happened.  `--quick`: the small correctness cases only."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdo_ptq_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
SMALL = [  # name, N, Cin, H, W, Cout, k, stride, transposed
    ("tiny 2 tiles", 1, 32, 16, 16, 32, 3, 1, False), ("tiny 64ch", 2, 64, 16, 16, 64, 3, 1, False),
    ("odd tiles", 1, 64, 24, 16, 96, 3, 1, False), ("ragged", 2, 96, 30, 44, 160, 3, 1, False),
    ("h_a.2 b8p", 8, 192, 16, 16, 192, 5, 2, False), ("g_s.2 b8p", 8, 192, 32, 32, 192, 5, 2, True),
    ("deconv ragged", 1, 64, 9, 13, 96, 5, 2, True),
]
BIG = [
    ("g_a.4 b8p", 8, 192, 64, 64, 192, 5, 2, False), ("g_a.2 b8p", 8, 192, 128, 128, 192, 5, 2, False),
    ("g_a.6 b8p", 8, 192, 32, 32, 320, 5, 2, False), ("g_s.4 b8p", 8, 192, 64, 64, 192, 5, 2, True),
    ("h_s.4 b8p", 8, 480, 16, 16, 640, 3, 1, False),
    ("g_a.2 2K", 1, 192, 768, 1024, 192, 5, 2, False), ("g_a.4 2K", 1, 192, 384, 512, 192, 5, 2, False),
    ("g_s.4 2K", 1, 192, 384, 512, 192, 5, 2, True), ("g_s.2 2K", 1, 192, 192, 256, 192, 5, 2, True),
]
cases = SMALL if "--quick" in sys.argv else SMALL + BIG
flush = torch.empty(64 * 1024 * 1024, device=dev)
lib = _lib.lib()
for name, N, Cin, H, W, Cout, k, st, tr in cases:
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    if tr:
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), st, k // 2, st - 1)
    else:
        ref = F.conv2d(x.double(), w.double(), b.double(), st, k // 2)
    delta, zp = ops.wq_init_minmax(w, 1 if tr else 0, 8)
    n_int = ops.wq_int_weights(w, delta, zp, 1 if tr else 0, 256)
    wq = (n_int * delta).double()
    ref_int = (F.conv_transpose2d(x.double(), wq, b.double(), st, k // 2, st - 1) if tr
               else F.conv2d(x.double(), wq, b.double(), st, k // 2))
    row = {"case": name}
    for sk in (1, 0):
        for mode in (0, 2):
            lib.b200lic_set_option(b"pair", mode)
            lib.b200lic_set_option(b"streamk", sk)
            tag = f"pair{mode}_sk{sk}"
            print(json.dumps({"case": name, "start": tag}), flush=True)
            d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0)
            fn = (lambda: ops.deconv2d_raw(x, w, b, d)) if tr else (lambda: ops.conv2d_raw(x, w, b, d))
            y = fn()
            y_int = ops.conv_wq(x, n_int, delta.reshape(-1).contiguous(), b, stride=st, padding=k // 2,
                                output_padding=st - 1 if tr else 0, transposed=tr)
            torch.cuda.synchronize()
            row[tag + "_err"] = float(f"{((y.double() - ref).norm() / ref.norm()).item():.2e}")
            if y_int is not None:
                row[tag + "_err_int"] = float(f"{((y_int.double() - ref_int).norm() / ref_int.norm()).item():.2e}")
            if "--quick" in sys.argv:
                continue
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
            row[tag + "_us"] = round(sorted(ts)[len(ts) // 2], 1)
    print(json.dumps(row), flush=True)
lib.b200lic_set_option(b"pair", 1)
lib.b200lic_set_option(b"streamk", 1)
