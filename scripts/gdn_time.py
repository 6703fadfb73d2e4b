import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
from rdo_ptq_b200 import ops
dev = "cuda"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for shape in [(1, 192, 768, 1024), (1, 192, 384, 512), (1, 192, 384, 256), (8, 192, 128, 128)]:
    N, Cc, H, W = shape
    x = torch.randn(shape, device=dev)
    gam = (torch.rand(Cc, Cc, device=dev) * 0.02 + 0.1 * torch.eye(Cc, device=dev))
    bet = 1 + torch.rand(Cc, device=dev)
    d = ops.gdn_desc(x.shape, False)
    packed = ops.pack_weights(gam.view(Cc, Cc, 1, 1), d, False)
    keys = ops.act_quant_stats(x)
    y = torch.empty_like(x)
    ws = ops._workspace(d, ops.fwd_op(False), x.device)
    slot = ops.conv_x_slot(d, False, ws)
    xq = torch.empty_like(x)
    def old():
        ops.act_quant_apply_stage(x, keys, 8, slot, square=True, out=xq)
        ops.conv_fwd_packed(None, packed, d, False, bias=bet, gdn_x=xq, ws=ws, y=y)
    def old_noq():
        ops.conv_fwd_packed(x, packed, d, False, bias=bet, gdn_x=x, ws=ws, y=y)
    el = x.numel()
    for name, fn in [("old+q", old), ("old", old_noq), ("new+q", lambda: ops.gdn_fwd_fused(x, packed, bet, False, pending=(keys, 8), y=y)),
                     ("new", lambda: ops.gdn_fwd_fused(x, packed, bet, False, y=y))]:
        us = t(fn)
        print(shape, name, f"{us:8.1f} us  {8 * el / us / 1e6:6.2f} TB/s algorithmic")
