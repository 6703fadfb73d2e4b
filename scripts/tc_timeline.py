"""B200LIC_TC_DEBUG=3 python scripts/tc_timeline.py [gdn|conv|deconv]  -- prints CTA 0's first-item timeline (ns)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rdo_ptq_b200 import ops, _lib
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "gdn"
g = torch.Generator().manual_seed(1)
B = 8
if what == "gdn":
    x = torch.randn(B, 192, 128, 128, generator=g).to(dev)
    gam = (torch.rand(192, 192, generator=g) * 0.01 + 0.1 * torch.eye(192)).to(dev)
    bet = torch.ones(192).to(dev)
    d = ops.gdn_desc(x.shape, False)
    fn = lambda: ops.conv2d_raw(x, gam.view(192, 192, 1, 1), bet, d, gdn_x=x)
elif what == "conv":
    x = torch.randn(B, 192, 128, 128, generator=g).to(dev)
    w = (torch.randn(192, 192, 5, 5, generator=g) * 0.05).to(dev)
    b = torch.randn(192, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, 2, 2)
    fn = lambda: ops.conv2d_raw(x, w, b, d)
elif what == "small":          # h_a.4 at batch 1 (768x512 image): one pixel tile, 150 K blocks -- the latency regime
    x = torch.randn(1, 192, 16, 24, generator=g).to(dev)
    w = (torch.randn(192, 192, 5, 5, generator=g) * 0.05).to(dev)
    b = torch.randn(192, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, 2, 2)
    fn = lambda: ops.conv2d_raw(x, w, b, d)
else:
    x = torch.randn(B, 192, 64, 64, generator=g).to(dev)
    w = (torch.randn(192, 192, 5, 5, generator=g) * 0.05).to(dev)
    b = torch.randn(192, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, 2, 2, True, 1)
    fn = lambda: ops.deconv2d_raw(x, w, b, d)
for _ in range(3):
    fn()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 128)()
n = _lib.lib().b200lic_debug_timeline(buf, 128)
t = list(buf)
t0 = t[0]
print("stamps relative to kernel start (us):")
print(" first stage landed (MMA thread):", (t[100] - t0) / 1e3, " last MMA issued:", (t[101] - t0) / 1e3)
print(" accumulators ready (epilogue):", (t[1] - t0) / 1e3, " item done:", (t[99] - t0) / 1e3)
for ci in range(30):
    a, b_, c = t[2 + 3 * ci], t[3 + 3 * ci], t[4 + 3 * ci]
    if a == 0:
        break
    print(f" chunk {ci:2d}: staging free {(a - t0) / 1e3:8.2f}  x landed {(b_ - t0) / 1e3:8.2f}  store issued {(c - t0) / 1e3:8.2f}")
if t[110]:
    names = ["tmem ld done", "bias added", "math done", "staged"]
    print(" chunk 1 in detail (us after its x landed):")
    base = t[3 + 3]
    for h in range(2):
        print("  half", h, " ".join(f"{names[k]} {(t[110 + 4 * h + k] - base) / 1e3:6.2f}" for k in range(4) if t[110 + 4 * h + k]))
    print(f"  proxy fence done {(t[118] - base) / 1e3:6.2f}  barrier passed {(t[119] - base) / 1e3:6.2f}")
