"""Launches each hot-path kernel a few times at the BASELINE shapes -- the command ncu wraps (see profiles/README.md).

  ncu --set full --clock-control none --import-source on -o gpurun_out/prof python scripts/profile_kernels.py --reps 1
  python scripts/profile_kernels.py --reps 20 --time        # CUDA-event timings + roofline fractions (no profiler)
  python scripts/profile_kernels.py --reps 20 --time --graph   # the same from a CUDA-graph replay of each case
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdo_ptq_b200 import ops, _lib  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="time a CUDA-graph replay of each case (device time of its kernels: no host launch gaps)")
    ap.add_argument("--only", default="")
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1005)
    B = args.batch
    hbm, tf = peaks()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    cases = []

    def conv_case(name, N, Cin, H, W, Cout, k, st, transposed=False):
        x = torch.randn(N, Cin, H, W, generator=g).to(dev)
        wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
        w = (torch.randn(wshape, generator=g) * 0.05).to(dev)
        b = torch.randn(Cout, generator=g).to(dev)
        d = ops.conv_desc(x.shape, w.shape, st, k // 2, transposed, st - 1 if transposed else 0)
        macs = (H * W if transposed else d.Ho * d.Wo) * N * Cin * Cout * k * k
        dy = torch.randn(N, Cout, d.Ho, d.Wo, generator=g).to(dev)
        dw = torch.empty_like(w)
        dx = torch.empty_like(x)
        fwd = (lambda: ops.deconv2d_raw(x, w, b, d)) if transposed else (lambda: ops.conv2d_raw(x, w, b, d))
        opw = _lib.OP_DECONV_WGRAD if transposed else _lib.OP_CONV_WGRAD
        opd = _lib.OP_DECONV_DGRAD if transposed else _lib.OP_CONV_DGRAD
        wsw, nw = ops._workspace(d, opw, dev)
        wsd, nd = ops._workspace(d, opd, dev)
        nm = "deconv" if transposed else "conv"
        cases.append((f"{name}.fwd", fwd, 2.0 * macs, "flop"))
        cases.append((f"{name}.wgrad", lambda: ops.call(nm + "_wgrad", C.byref(d), ops._p(x), ops._p(dy), ops._p(dw),
                                                        ops._p(wsw), nw), 2.0 * macs, "flop"))
        cases.append((f"{name}.dgrad", lambda: ops.call(nm + "_dgrad", C.byref(d), ops._p(dy), ops._p(w), ops._p(dx),
                                                        ops._p(wsd), nd), 2.0 * macs, "flop"))

    conv_case("g_a.2 conv5x5s2 192->192 @128x128", B, 192, 128, 128, 192, 5, 2)
    conv_case("g_s.4 deconv5x5s2 192->192 @64x64", B, 192, 64, 64, 192, 5, 2, True)
    conv_case("h_s.4 conv3x3s1 480->640 @16x16", B, 480, 16, 16, 640, 3, 1)

    # GDN 192 @128x128
    xg = torch.randn(B, 192, 128, 128, generator=g).to(dev)
    gam = (torch.rand(192, 192, generator=g) * 0.01 + 0.1 * torch.eye(192)).to(dev)
    bet = torch.ones(192).to(dev)
    dg = ops.gdn_desc(xg.shape, False)
    cases.append(("gdn 192 @128x128", lambda: ops.conv2d_raw(xg, gam.view(192, 192, 1, 1), bet, dg, gdn_x=xg),
                  2.0 * B * 128 * 128 * 192 * 192, "flop"))

    # entropy / loss / quantiser kernels at a batch that leaves the launch-latency regime
    NB = 24
    y = (torch.randn(NB, 320, 96, 128, generator=g) * 3).to(dev)
    par = torch.randn(NB, 640, 96, 128, generator=g).to(dev)
    sc, mu = par.chunk(2, 1)
    cases.append((f"gaussian_lik [{NB},320,96,128] (no lik)", lambda: ops.gaussian_lik(y, sc, mu, want_lik=False),
                  16.0 * y.numel(), "byte"))
    z = (torch.randn(NB * 8, 192, 24, 32, generator=g) * 2).to(dev)
    pk = torch.randn(192, 58, generator=g).to(dev) * 0.5
    med = torch.zeros(192, device=dev)
    tab = ops.factorized_table(pk, med)
    cases.append((f"factorized_lik [{NB * 8},192,24,32] (no lik)",
                  lambda: ops.factorized_lik(z, pk, med, want_lik=False, table=tab), 8.0 * z.numel(), "byte"))
    cases.append((f"factorized_lik [{NB * 8},192,24,32] (no lik, tables built in the kernel)",
                  lambda: ops.factorized_lik(z, pk, med, want_lik=False), 8.0 * z.numel(), "byte"))
    a = torch.randn(B, 192, 128, 128, generator=g).to(dev)
    b2 = torch.randn(B, 192, 128, 128, generator=g).to(dev)
    cases.append(("lp_loss_fwd_bwd [8,192,128,128]", lambda: ops.lp_loss_fwd_bwd(a, b2), 12.0 * a.numel(), "byte"))
    cases.append(("sq_err_sum [8,192,128,128]", lambda: ops.sq_err_sum(a, b2), 8.0 * a.numel(), "byte"))
    cases.append(("act_quant [8,192,128,128] (stats+apply)", lambda: ops.act_quant(a), 12.0 * a.numel(), "byte"))
    idx = torch.arange(B, device=dev)
    cases.append(("gather_mix [8,192,128,128]", lambda: ops.gather_mix(a, b2, idx, prob=0.5, seed=1), 12.0 * a.numel(),
                  "byte"))
    w = (torch.randn(640, 480, 3, 3, generator=g) * 0.05).to(dev)
    dl, zp = ops.wq_init_minmax(w, 0)
    al = ops.adaround_init_alpha(w, dl, 0)
    m1, m2 = torch.zeros_like(w), torch.zeros_like(w)
    dwq = torch.randn_like(w)
    cases.append(("wq_init_minmax [640,480,3,3]", lambda: ops.wq_init_minmax(w, 0), 4.0 * w.numel(), "byte"))
    cases.append(("adaround_fwd [640,480,3,3]", lambda: ops.adaround_fwd(w, al, dl, zp, 0, 256, True), 12.0 * w.numel(),
                  "byte"))
    cases.append(("adaround_bwd_adam [640,480,3,3]",
                  lambda: ops.adaround_bwd_adam(w, al, dl, zp, dwq, m1, m2, 0, 256, 5, reg_weight=0.01, reg_b=10.0),
                  32.0 * w.numel(), "byte"))

    rows = []
    flush_ms = [None]
    for name, fn, work, kind in cases:
        if args.only and args.only not in name:
            continue
        if not args.time:
            for _ in range(args.reps):
                fn()
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            fn()
        if args.graph:
            # device time of the case's kernels with a cold L2: graph A = R x (flush, case), graph B = R x (flush);
            # (A - B) / R.  Eager issue would time the host (allocations, ctypes, launch gaps) for the short kernels.
            R = 5
            torch.cuda.synchronize()

            def timed(body):
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    for _ in range(R):
                        body()
                out = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    cg.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    out.append(e0.elapsed_time(e1))
                out.sort()
                return out[len(out) // 2]

            def both():
                flush.zero_()
                fn()
            if flush_ms[0] is None:
                flush_ms[0] = timed(flush.zero_)
            ts = [(timed(both) - flush_ms[0]) / R]
        else:
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        if kind == "flop":
            ach, pk_, unit = work / ms / 1e9, tf, "TFLOP/s"
        else:
            ach, pk_, unit = work / ms / 1e6, hbm, "GB/s"
        rows.append(dict(kernel=name, ms=round(ms, 4), achieved=round(ach, 1), unit=unit, peak=pk_,
                         frac=round(ach / pk_, 3)))
        print(json.dumps(rows[-1]))
    if args.time:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "kernel_times_graph.json" if args.graph else "kernel_times.json"
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)


if __name__ == "__main__":
    main()
