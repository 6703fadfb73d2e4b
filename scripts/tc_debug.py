"""GPU debug: tcgen05 engine vs the SIMT engine (and torch CPU for small cases) on selected shapes, with timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from rdo_ptq_b200 import ops
from rdo_ptq_b200._lib import ENGINE_SIMT, ENGINE_TC

dev = torch.device("cuda:0")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def conv_case(N, Cin, H, W, Cout, k, st, pd, cpu_check=False):
    g = torch.Generator().manual_seed(N + Cin + H + Cout + k)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ds = ops.conv_desc(x.shape, w.shape, st, pd, engine=ENGINE_SIMT, act=ops.ACT_LEAKY_RELU)
    dt = ops.conv_desc(x.shape, w.shape, st, pd, engine=ENGINE_TC, act=ops.ACT_LEAKY_RELU)
    ys = ops.conv2d_raw(x, w, b, ds)
    yt = ops.conv2d_raw(x, w, b, dt)
    torch.cuda.synchronize()
    msg = f"conv N{N} {Cin}->{Cout} {H}x{W} k{k} s{st}: tc-vs-simt rel {rel(yt, ys):.2e}"
    if cpu_check:
        ref = F.leaky_relu(F.conv2d(x.cpu(), w.cpu(), b.cpu(), stride=st, padding=pd), 0.01)
        msg += f" | simt-vs-cpu {rel(ys.cpu(), ref):.2e} tc-vs-cpu {rel(yt.cpu(), ref):.2e}"
    macs = N * ys.shape[2] * ys.shape[3] * Cout * Cin * k * k
    ts, tt = timeit(lambda: ops.conv2d_raw(x, w, b, ds)), timeit(lambda: ops.conv2d_raw(x, w, b, dt))
    msg += f" | simt {ts:.3f} ms ({2*macs/ts/1e9:.1f} TF/s) tc {tt:.3f} ms ({2*macs/tt/1e9:.1f} TF/s)"
    print(msg, flush=True)


def deconv_case(N, Cin, H, W, Cout, k, st, pd, op, cpu_check=False):
    g = torch.Generator().manual_seed(N + Cin + H + Cout + k)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cin, Cout, k, k, generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ds = ops.conv_desc(x.shape, w.shape, st, pd, True, op, engine=ENGINE_SIMT)
    dt = ops.conv_desc(x.shape, w.shape, st, pd, True, op, engine=ENGINE_TC)
    ys = ops.deconv2d_raw(x, w, b, ds)
    yt = ops.deconv2d_raw(x, w, b, dt)
    torch.cuda.synchronize()
    msg = f"deconv N{N} {Cin}->{Cout} {H}x{W} k{k} s{st}: tc-vs-simt rel {rel(yt, ys):.2e}"
    if cpu_check:
        ref = F.conv_transpose2d(x.cpu(), w.cpu(), b.cpu(), stride=st, padding=pd, output_padding=op)
        msg += f" | tc-vs-cpu {rel(yt.cpu(), ref):.2e}"
    macs = N * H * W * Cout * Cin * k * k
    ts, tt = timeit(lambda: ops.deconv2d_raw(x, w, b, ds)), timeit(lambda: ops.deconv2d_raw(x, w, b, dt))
    msg += f" | simt {ts:.3f} ms ({2*macs/ts/1e9:.1f} TF/s) tc {tt:.3f} ms ({2*macs/tt/1e9:.1f} TF/s)"
    print(msg, flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    conv_case(1, 64, 16, 16, 64, 1, 1, 0, True)
    conv_case(1, 64, 16, 16, 64, 3, 1, 1, True)
    conv_case(2, 48, 20, 28, 80, 3, 2, 1, True)
    conv_case(1, 192, 32, 48, 192, 5, 2, 2, True)
    deconv_case(1, 64, 8, 8, 64, 3, 1, 1, 0, True)
    deconv_case(2, 192, 16, 24, 192, 5, 2, 2, 1, True)
    deconv_case(1, 320, 8, 12, 192, 5, 2, 2, 1, True)
    if which == "all":
        conv_case(8, 192, 128, 128, 192, 5, 2, 2)
        conv_case(8, 192, 32, 32, 320, 5, 2, 2)
        conv_case(8, 192, 64, 64, 192, 1, 1, 0)
        deconv_case(8, 192, 64, 64, 192, 5, 2, 2, 1)


def wgrad_case(N, Cin, H, W, Cout, k, st, pd, transposed=False, op=0, cpu_check=False):
    import ctypes as C
    from rdo_ptq_b200 import _lib
    g = torch.Generator().manual_seed(N + Cin + H + Cout + k)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
    ds = ops.conv_desc(x.shape, wshape, st, pd, transposed, op, engine=ENGINE_SIMT)
    dt = ops.conv_desc(x.shape, wshape, st, pd, transposed, op, engine=ENGINE_TC)
    dy = torch.randn(N, Cout, ds.Ho, ds.Wo, generator=g).to(dev)
    name = "deconv_wgrad" if transposed else "conv_wgrad"
    opid = _lib.OP_DECONV_WGRAD if transposed else _lib.OP_CONV_WGRAD

    def run(d):
        dw = torch.empty(wshape, device=dev)
        ws, nws = ops._workspace(d, opid, dev)
        ops.call(name, C.byref(d), ops._p(x), ops._p(dy), ops._p(dw), ops._p(ws), nws)
        return dw
    a, b = run(ds), run(dt)
    torch.cuda.synchronize()
    msg = f"{name} N{N} {Cin}->{Cout} {H}x{W} k{k} s{st}: tc-vs-simt rel {rel(b, a):.2e}"
    if cpu_check:
        xc = x.cpu().requires_grad_(False)
        wc = torch.zeros(wshape, requires_grad=True)
        y = (F.conv_transpose2d(xc, wc, None, st, pd, op) if transposed else F.conv2d(xc, wc, None, st, pd))
        y.backward(dy.cpu())
        msg += f" | tc-vs-cpu {rel(b.cpu(), wc.grad):.2e}"
    macs = N * (H * W if transposed else ds.Ho * ds.Wo) * Cout * Cin * k * k
    ts, tt = timeit(lambda: run(ds)), timeit(lambda: run(dt))
    msg += f" | simt {ts:.3f} ms ({2*macs/ts/1e9:.1f} TF/s) tc {tt:.3f} ms ({2*macs/tt/1e9:.1f} TF/s)"
    print(msg, flush=True)


if __name__ == "__main__":
    wgrad_case(1, 64, 16, 16, 64, 1, 1, 0, cpu_check=True)
    wgrad_case(2, 64, 16, 16, 128, 3, 1, 1, cpu_check=True)
    wgrad_case(2, 48, 20, 28, 80, 3, 2, 1, cpu_check=True)
    wgrad_case(1, 192, 32, 48, 192, 5, 2, 2, cpu_check=True)
    wgrad_case(2, 192, 8, 12, 320, 5, 2, 2, True, 1, cpu_check=True)
    if (sys.argv[1] if len(sys.argv) > 1 else "all") == "all":
        wgrad_case(8, 192, 128, 128, 192, 5, 2, 2)
        wgrad_case(8, 192, 64, 64, 192, 5, 2, 2, True, 1)
        wgrad_case(8, 192, 64, 64, 192, 1, 1, 0)
