import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import torch.nn.functional as F
from rdo_ptq_b200 import ops, _lib
from rdo_ptq_b200._lib import ENGINE_SIMT, ENGINE_TC
dev = torch.device("cuda:0")
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
N, Cin, H, W, Cout, k, st, pd, op = 1, 192, 8, 12, 192, 5, 2, 2, 1
g = torch.Generator().manual_seed(sum((N, Cin, H, W, Cout, k, st, pd, op)))
x = torch.randn(N, Cin, H, W, generator=g)
w = torch.randn(Cin, Cout, k, k, generator=g) * 0.1
b = torch.randn(Cout, generator=g)
dy = torch.randn(N, Cout, 16, 24, generator=g)
for scale in (1.0, 0.01):
    dyd = (dy * scale).to(dev)
    res = {}
    for eng in (ENGINE_SIMT, ENGINE_TC):
        d = ops.conv_desc(x.shape, w.shape, st, pd, True, op, engine=eng)
        outs = []
        for rep in range(4):
            dx = torch.empty(N, Cin, H, W, device=dev)
            ws, nws = ops._workspace(d, _lib.OP_DECONV_DGRAD, dev)
            ops.call("deconv_dgrad", C.byref(d), ops._p(dyd), ops._p(w.to(dev)), ops._p(dx), ops._p(ws), nws)
            torch.cuda.synchronize()
            outs.append(dx.clone())
        res[eng] = outs
        print("scale", scale, "engine", eng, "self-consistency", [rel(o, outs[0]) for o in outs[1:]])
    print("scale", scale, "tc vs simt", rel(res[ENGINE_TC][0], res[ENGINE_SIMT][0]))
    ref = F.conv2d(dy * scale, w, None, stride=st, padding=pd)
    print("  simt vs cpu", rel(res[ENGINE_SIMT][0].cpu(), ref), " tc vs cpu", rel(res[ENGINE_TC][0].cpu(), ref))
# the same through the autograd path of the test
xd, wd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
out = ops.conv_transpose2d(xd, wd, b.to(dev), st, pd, op, act=ops.ACT_LEAKY_RELU, slope=0.01)
out.backward(dy.to(dev))
xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
refo = F.leaky_relu(F.conv_transpose2d(xr, wr, b, stride=st, padding=pd, output_padding=op), 0.01)
refo.backward(dy)
print("autograd path: fwd", rel(out.cpu(), refo), "dx", rel(xd.grad.cpu(), xr.grad), "dw", rel(wd.grad.cpu(), wr.grad))
dyeff = ops.act_bwd(out.detach(), dy.to(dev), ops.ACT_LEAKY_RELU, 0.01)
mask_ref = torch.where(refo > 0, dy, dy * 0.01)
print("act_bwd vs cpu mask", rel(dyeff.cpu(), mask_ref), "sign flips", ((out.cpu() > 0) != (refo > 0)).sum().item())
