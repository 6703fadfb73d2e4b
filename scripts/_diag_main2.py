import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import codec as ocodec, quant_wrap as owrap, calib as ocalib, evalpath as oeval
from rdo_ptq_b200 import synth, main2, evaluate as E
import test_gpu_model as T
dev = torch.device("cuda:0")
args = main2.parse_args(["--arch", "Minnen2018", "--n_bits_w", "4", "--channel_wise", "--batch_size", "2", "--num_samples", "4", "--iters_w", "12", "--test_before_calibration"])
om, pm, Q = T.build_pair("mbt2018-mean", dict(N=16, M=24), 1.2, dev)
cali = synth.calibration_patches(4, 64)
imgs = synth.synthetic_images(2, 100, 150)
with torch.no_grad():
    om(cali[:1])
wq = {'n_bits': 4, 'channel_wise': True, 'scale_method': 'max'}
aq = {'n_bits': 8, 'channel_wise': True, 'scale_method': 'max', 'leaf_param': False}
oqm = owrap.QuantModel(om, wq, aq).eval()
oqm.set_first_last_layer_to_8bit()
oqm.disable_network_output_quantization()
oqm.set_quant_state(True, False)
with torch.no_grad():
    oqm(cali[:2])
print("oracle nearest:", oeval.evaluate(oqm, imgs))
oqm.model.g_s[-1].set_quant_state(True, False)
otr = ocalib.recon_model(oqm, cali, batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5, plan=ocalib.DrawPlan())
oqm.set_quant_state(True, False)
print("oracle w_opt:", oeval.evaluate(oqm, imgs))
pqm, rep = main2.optimize_model(args, model=pm, cali_data=cali, test_images=imgs, device=dev, plan=T.ReplayPlan())
for k in ("fp32", "w_nearest", "w_opt", "wa_opt"):
    print("product", k, rep[k]["per_image"])
pqm.set_quant_state(True, False)
for graph in (False, True):
    print("product w_opt graph=", graph, E.evaluate(pqm, [i.to(dev) for i in imgs], shard=False, graph=graph)["per_image"])
x = oeval.pad(imgs[0], 256)
ref, rows = T.per_layer_io(oqm, x, (owrap.QuantModule,))
pmods = dict((n, m) for n, m in pqm.named_modules() if isinstance(m, Q.QuantModule))
for name, xi, yo in rows:
    with torch.no_grad():
        out = pmods[name](xi.to(dev))
    print(name, "layer-local rel err", T.rel_err(out, yo))
with torch.no_grad():
    out = pqm(x.to(dev))
print("x_hat rel err", T.rel_err(out["x_hat"], ref["x_hat"]), "lik_y", T.rel_err(out["likelihoods"]["y"], ref["likelihoods"]["y"]),
      "lik_z", T.rel_err(out["likelihoods"]["z"], ref["likelihoods"]["z"]))
print("bpp", E.compute_bpp(out), oeval.compute_bpp(ref))
