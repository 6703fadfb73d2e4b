"""Calibration-sweep throughput (bench.py's `value` definition: units * batch / step time, caches resident in HBM) for any
of the three codecs -- BASELINE configs 2 (mbt2018-mean) and 3 (cheng2020-attn: residual blocks reconstructed jointly,
masked context model evaluated in parallel).

  python scripts/calib_throughput.py --arch cheng2020-attn --arch mbt2018-mean --arch bmshj2018-hyperprior
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from rdo_ptq_b200 import codec, synth  # noqa: E402
from rdo_ptq_b200.quantization import QuantModel  # noqa: E402
from rdo_ptq_b200.quantization.session import CalibrationSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", action="append")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--pool", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for arch in args.arch or ["cheng2020-attn"]:
        torch.manual_seed(1005)
        kw = dict(N=192) if arch == "cheng2020-attn" else dict(N=192, M=320)
        m = codec.ARCHS[arch](**kw).eval()
        synth.init_weights(m, gain=0.6 if arch == "cheng2020-attn" else 1.2)
        m.to(dev)
        cali = synth.calibration_patches(args.pool, 256, seed=1005).to(dev)
        with torch.no_grad():
            m(cali[:1])                                  # one FP forward first: bakes the MaskedConv2d mask
        qnn = QuantModel(m, B.WQ, B.AQ, is_cheng=(arch == "cheng2020-attn")).eval()
        sess = CalibrationSession(qnn, cali, batch_size=args.batch, **B.CALIB)
        for _ in range(4):
            sess.sweep()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            sess.sweep()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        losses = sess.losses()
        ok = all(v["rec"] == v["rec"] for v in losses.values())
        print(json.dumps({"arch": arch, "units": len(sess.units), "batch": args.batch, "ms_per_step": ms,
                          "calib_imgs_s": len(sess.units) * args.batch / (ms / 1e3), "losses_finite": ok}))
        del sess, qnn, m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
