"""W8A8 evaluation forward (BASELINE configs 4/5) of one synthetic image: the command ncu wraps for the launch list
of the forward, and a CUDA-event timer for it (graph replay, like evaluate.evaluate).

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fwd_launches.csv \
      python scripts/profile_fwd.py --eager --reps 1
  python scripts/profile_fwd.py --hw 512x768 --hw 1365x2048 --arch mbt2018-mean --arch cheng2020-attn
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdo_ptq_b200 import codec, synth, _lib, evaluate as E  # noqa: E402
from rdo_ptq_b200.quantization import QuantModel  # noqa: E402

WQ = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)


def build(arch, dev, n_bits=8):
    torch.manual_seed(1005)
    kw = dict(N=192) if arch == "cheng2020-attn" else dict(N=192, M=320)
    m = codec.ARCHS[arch](**kw).eval()
    synth.init_weights(m, gain=1.2)
    m.to(dev)
    wq, aq = dict(WQ, n_bits=n_bits), dict(AQ, n_bits=n_bits)
    return QuantModel(m, wq, aq, is_cheng=(arch == "cheng2020-attn")).eval()


def w8a8(qnn, x):
    """main2.py:272-282: weights quantised (nearest here: no calibration), dynamic A8 on, image output layer A-off."""
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(x)
    for m in qnn.modules():
        if hasattr(m, "trained"):
            m.trained = True
    qnn.set_quant_state(True, True)
    last = qnn.model.g_s[-1]
    if hasattr(last, "set_quant_state"):
        last.set_quant_state(True, False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", action="append")
    ap.add_argument("--hw", action="append")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--bits", type=int, default=8, help="W/A bit width (10 = BASELINE config 4, W10A10)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for arch in args.arch or ["mbt2018-mean"]:
        from rdo_ptq_b200.quantization.quantizer import UniformAffineQuantizer
        UniformAffineQuantizer.act_bits_follow_n_bits = args.bits != 8
        qnn = build(arch, dev, args.bits)
        first = True
        for hw in args.hw or ["512x768"]:
            h, w = (int(v) for v in hw.split("x"))
            x = E.pad(synth.synthetic_image(h, w).to(dev), 256)
            if first:
                w8a8(qnn, x)
                first = False
            with torch.no_grad():
                if args.eager:
                    from rdo_ptq_b200 import ops
                    with ops.defer_actq():                  # what evaluate.GraphedForward captures
                        qnn(x)
                        qnn(x)
                        torch.cuda.synchronize()
                        n0 = _lib.launch_count()
                        torch.cuda.profiler.start()         # ncu --profile-from-start off: the steady-state forward only
                        for _ in range(args.reps):
                            out = qnn(x)
                            E.total_bits(out)
                        torch.cuda.synchronize()
                        torch.cuda.profiler.stop()
                    print(json.dumps({"arch": arch, "hw": hw, "launches_per_fwd": (_lib.launch_count() - n0) / args.reps}))
                    continue
                gf = E.GraphedForward(qnn)
                for _ in range(3):
                    gf(x)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(args.reps):
                    gf(x)
                b.record()
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.reps
            print(json.dumps({"arch": arch, "bits": args.bits, "hw": hw, "padded": list(x.shape[2:]), "ms_per_image": ms,
                              "mpx_s": h * w / 1e6 / (ms / 1e3)}))


if __name__ == "__main__":
    main()
