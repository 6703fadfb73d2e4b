"""GPU: true-integer deployment: a calibrated QuantModel exported as integer codes + per-channel
(delta, zero_point) and loaded back reproduces the calibrated model's quantised forward BIT FOR BIT, from a file a quarter
of the size -- the round trip light-uniform-PTQ's INT8.pth cannot make (LU README.md:93-95, quantize.py:155-157)."""
import io
import math

import pytest
import torch

from rdo_ptq_b200 import synth

pytestmark = pytest.mark.gpu


def _same_forward(a, b, x):
    with torch.no_grad():
        oa, ob = a(x), b(x)
    assert torch.equal(oa["x_hat"], ob["x_hat"])
    assert torch.equal(oa["likelihoods"]["y"], ob["likelihoods"]["y"])
    assert torch.equal(oa["likelihoods"]["z"], ob["likelihoods"]["z"])


def test_export_load_roundtrip_after_adaround_calibration(dev, tmp_path):
    from rdo_ptq_b200 import deploy, main2
    args = main2.parse_args(["--arch", "mbt2018-mean", "--N", "16", "--M", "24", "--n_bits_w", "4", "--channel_wise",
                             "--batch_size", "2", "--num_samples", "4", "--iters_w", "8", "--patch", "64",
                             "--test_hw", "64x96", "--n_test", "1"])
    qnn, rep = main2.optimize_model(args, device=dev)
    kw = dict(N=16, M=24)
    path = tmp_path / "mbt_w4.b200int8"
    blob = deploy.export_int8(qnn, "mbt2018-mean", kw, path)
    first, second = blob["layers"]["g_a.0"], blob["layers"]["g_a.2"]
    assert first["n_bits"] == 8 and second["n_bits"] == 4                    # 8-bit head / stem survives the export
    assert second["codes"].dtype == torch.uint8 and int(second["codes"].max()) <= 15
    assert all(L["codes"].dtype == torch.uint8 for L in blob["layers"].values()) and len(blob["layers"]) == 20
    buf = io.BytesIO()
    torch.save(qnn.model.state_dict(), buf)
    assert path.stat().st_size < 0.35 * buf.getbuffer().nbytes               # codes are bytes; no alpha, no fp32 copy
    back = deploy.load_int8(str(path), device=dev)
    x = synth.synthetic_image(64, 96).to(dev)
    qnn.eval()
    for w_a in ((True, False), (True, True)):
        for q in (qnn, back):
            q.set_quant_state(*w_a)
            q.model.g_s[-1].set_quant_state(True, False)
        _same_forward(qnn, back, x)
    # the loaded model's codes are the exported codes
    for name, m in back.model.named_modules():
        if name in blob["layers"]:
            assert torch.equal(m.weight_quantizer.codes(m.weight).cpu().to(torch.uint8), blob["layers"][name]["codes"])
    assert math.isfinite(rep["wa_opt"]["bpp"])


def test_light_uniform_int8_reloads(dev, tmp_path):
    """The INT8 model light-uniform-PTQ cannot reload: uint8 codes + delta / zero_point travel together here."""
    from rdo_ptq_b200 import deploy, quantize as lu_entry, quant_int as LU
    args = lu_entry.parse_args(["--N", "8", "--M", "12", "--hw", "64x96", "--n_test", "1"])
    qnn, _ = lu_entry.quantize_int8(args, device=dev)
    path = tmp_path / "lu.b200int8"
    blob = deploy.export_int8(qnn, args.arch, dict(N=8, M=12), path)
    assert blob["rules"] == "LU" and all(L["codes"].dtype == torch.uint8 for L in blob["layers"].values())
    back = deploy.load_int8(str(path), device=dev)
    assert isinstance(back, LU.QuantModel)
    back.set_quant_state(True, True)
    qnn.set_quant_state(True, True)
    assert [m.disable_act_quant for m in back.model.modules() if isinstance(m, LU.QuantModule)] == \
        [m.disable_act_quant for m in qnn.model.modules() if isinstance(m, LU.QuantModule)]
    _same_forward(qnn, back, synth.synthetic_image(64, 96).to(dev))


def test_export_refuses_an_uncalibrated_or_soft_model(dev):
    from rdo_ptq_b200 import codec, deploy, quantization as Q
    m = codec.ARCHS["mbt2018-mean"](N=8, M=12).eval().to(dev)
    q = Q.QuantModel(m, dict(n_bits=8, channel_wise=True, scale_method="max"),
                     dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)).eval()
    with pytest.raises(RuntimeError):
        deploy.export_int8(q, "mbt2018-mean", dict(N=8, M=12))
