"""Host-side issue time vs device time of CalibrationSession.sweep() at the bench workload (is the sweep launch-bound?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from rdo_ptq_b200 import codec, synth
from rdo_ptq_b200.quantization import QuantModel
from rdo_ptq_b200.quantization.session import CalibrationSession
dev = torch.device("cuda:0")
torch.manual_seed(1005)
m = codec.ARCHS[B.ARCH](N=B.N_CH, M=B.M_CH).eval()
synth.init_weights(m, gain=B.GAIN)
m.to(dev)
qnn = QuantModel(m, B.WQ, B.AQ).eval()
cali = synth.calibration_patches(B.POOL, B.PATCH, seed=1005).to(dev)
kw = {}
if len(sys.argv) > 1:
    kw["n_streams"] = int(sys.argv[1])
sess = CalibrationSession(qnn, cali, batch_size=B.PER_GPU_BATCH, **B.CALIB, **kw)
for _ in range(5):
    sess.sweep()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10):
        sess.sweep()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"streams={sess.n_streams}: host issue {1e3 * (t1 - t0) / 10:.3f} ms/sweep, total {1e3 * (t2 - t0) / 10:.3f} ms/sweep")
