"""One streaming-calibration step (bench.py's `e2e` leg: images H2D, two forwards that rebuild every unit's input /
target, the sweep, losses D2H) between cudaProfilerStart / Stop -- the command ncu wraps for the launch list of the step.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_e2e.csv python scripts/e2e_launches.py
"""
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
    dev = torch.device("cuda:0")
    torch.manual_seed(1005)
    m = codec.ARCHS[B.ARCH](N=B.N_CH, M=B.M_CH).eval()
    synth.init_weights(m, gain=B.GAIN)
    m.to(dev)
    cali = synth.calibration_patches(16, B.PATCH)
    with torch.no_grad():
        m(cali[:1].to(dev))
    qnn = QuantModel(m, B.WQ, B.AQ).eval()
    sess = CalibrationSession(qnn, cali, batch_size=B.PER_GPU_BATCH, host_caches="stream", **B.CALIB, **B.SEQUENTIAL)
    for _ in range(4):
        sess.sweep()
        sess.losses(lag=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    sess.sweep()
    sess.losses(lag=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
