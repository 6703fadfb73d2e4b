"""One calibration sweep (bench.py's headline workload, sequential schedule) between cudaProfilerStart / Stop: the command
ncu wraps for the launch list of the step.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_sweep.csv python scripts/sweep_launches.py [--fused 0] [--arch cheng2020-attn]
  python scripts/launch_summary.py gpurun_out/launches_sweep.csv            # per-kernel / per-unit table
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from rdo_ptq_b200 import codec, synth  # noqa: E402
from rdo_ptq_b200.quantization import QuantModel, recon  # noqa: E402
from rdo_ptq_b200.quantization.session import CalibrationSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--arch", default=B.ARCH)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--sweeps", type=int, default=1)
    args = ap.parse_args()
    recon.FUSED_DEFAULT = bool(args.fused)
    dev = torch.device("cuda:0")
    torch.manual_seed(1005)
    kw = dict(N=B.N_CH) if args.arch == "cheng2020-attn" else dict(N=B.N_CH, M=B.M_CH)
    m = codec.ARCHS[args.arch](**kw).eval()
    synth.init_weights(m, gain=0.6 if args.arch == "cheng2020-attn" else B.GAIN)
    m.to(dev)
    cali = synth.calibration_patches(args.pool, B.PATCH).to(dev)
    with torch.no_grad():
        m(cali[:1])
    qnn = QuantModel(m, B.WQ, B.AQ, is_cheng=(args.arch == "cheng2020-attn")).eval()
    sess = CalibrationSession(qnn, cali, batch_size=B.PER_GPU_BATCH, **B.CALIB, **B.SEQUENTIAL)
    for _ in range(4):
        sess.sweep()
    torch.cuda.synchronize()
    print("units:", [n for n, _ in sess.units])
    torch.cuda.profiler.start()
    for _ in range(args.sweeps):
        sess.sweep()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
