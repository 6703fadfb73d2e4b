"""CPU: the oracle's wrapper / calibration restatement (oracle.quant_wrap, oracle.calib) against the outputs of the
reference's OWN `task-oriented-PTQ/quantization` and `light-uniform-PTQ/quant_int` packages, stored in
tests/golden/wrap_ref.pt by oracle/make_golden.py::wrap_vectors (which imports the unmodified reference files through
oracle/_ref_shim.py).  Everything here is bit-exact (`torch.equal`) and runs without /root/reference."""
import copy
import os

import pytest
import torch

from oracle import calib as ocal, evalpath, make_golden as MG, quant_wrap as ow, quantizers as oq, _ref_shim

KINDS = (ow.QuantModule, ow.BaseQuantBlock)


def _build(case):
    fp, x = MG.build_fp_model(case["arch"], case["kw"], case["gain"])
    assert torch.equal(x, case["x"])
    q = ow.QuantModel(copy.deepcopy(fp), case["wq"], MG.AQ8, is_cheng=case["is_cheng"]).eval()
    if case["head8"]:
        q.set_first_last_layer_to_8bit()
    q.disable_network_output_quantization()
    return q, x


@pytest.mark.parametrize("arch", ["mbt2018-mean", "bmshj2018-hyperprior", "cheng2020-attn"])
@pytest.mark.parametrize("tag", ["w8", "w4"])
def test_quant_model_forwards_match_the_reference(golden_w, arch, tag):
    """QuantModel graph rewrite (quant_model.py:23-62), 8-bit head / stem (:81-98), QuantModule.forward
    (quant_layer.py:107-134) and the Cheng2020 blocks (quant_block.py:219-313): FP, W-only and W+A8 forwards."""
    case = golden_w[f"model/{arch}/{tag}"]
    q, x = _build(case)
    assert MG._structure(q, *KINDS) == case["structure"]
    assert [(n, m.weight_quantizer.n_bits, m.act_quantizer.n_bits, m.disable_act_quant)
            for n, m in q.named_modules() if isinstance(m, ow.QuantModule)] == case["n_bits"]
    for state in ("fp", "w", "wa"):
        if state == "wa":
            MG._mark_trained(q, KINDS)
        q.set_quant_state(state != "fp", state == "wa")
        if state == "wa":
            MG._output_layer(q, case["is_cheng"]).set_quant_state(True, False)
        out, layers = MG._layer_outputs(q, x, KINDS)
        ref = case[state]
        assert torch.equal(out["x_hat"], ref["x_hat"]), (arch, tag, state)
        assert torch.equal(out["likelihoods"]["y"], ref["lik_y"]) and torch.equal(out["likelihoods"]["z"], ref["lik_z"])
        for k, v in ref["layers"].items():
            assert torch.equal(layers[k], v), (arch, tag, state, k)
        assert evalpath.compute_bpp(out) == ref["bpp"]
    for n, m in q.named_modules():
        if isinstance(m, ow.QuantModule) and m.weight is not None:
            assert torch.equal(m.weight_quantizer.codes(m.weight), case["codes"][n]), n


@pytest.mark.parametrize("arch", ["mbt2018-mean", "cheng2020-attn"])
def test_calibration_walk_matches_the_reference(golden_w, arch):
    """save_inp_oup_data (utils.py:92-139), layer_reconstruction / block_reconstruction (layer_opt.py / block_opt.py
    :175-323) walked in main2.py:227-250 order: alpha and the hardened weights of every unit after the reference's own
    loop ran with the draws of oracle.calib.DrawPlan."""
    walk = golden_w[f"walk/{arch}"]
    c = walk["calib"]
    fp, x = MG.build_fp_model(walk["arch"], walk["kw"], walk["gain"])
    from rdo_ptq_b200 import synth
    cali = synth.calibration_patches(c["n_samples"], c["patch"])
    assert cali.double().sum() == walk["cali_sum"]
    q = ow.QuantModel(copy.deepcopy(fp), walk["wq"], MG.AQ8, is_cheng=walk["is_cheng"]).eval()
    q.set_first_last_layer_to_8bit()
    q.disable_network_output_quantization()
    q.set_quant_state(True, False)
    with torch.no_grad():
        q(cali[:c["batch_size"]])
    MG._output_layer(q, walk["is_cheng"]).set_quant_state(True, False)
    units = []

    def collect(module, prefix):
        for name, m in module.named_children():
            full = f"{prefix}.{name}" if prefix else name
            if isinstance(m, KINDS):
                units.append((full, name, m))
            else:
                collect(m, full)
    collect(q, "")
    assert [u[0] for u in units] == walk["units"]
    plan = ocal.DrawPlan(walk["seed"])
    seen = 0
    for uid, (full, name, m) in enumerate(units):
        if walk["limit"] is not None and uid >= walk["limit"]:
            break                                          # the reference walk stopped here (cheng2020-attn: 14 units)
        if full in walk["caches"]:
            (qi, fi), fo = ocal.save_inp_oup_data(q, m, cali, False, isinstance(m, ow.BaseQuantBlock))
            ref = walk["caches"][full]
            assert torch.equal(qi, ref["quant_in"]) and torch.equal(fi, ref["fp_in"]) and torch.equal(fo, ref["fp_out"])
        ocal.reconstruct(q, m, uid, name, cali, batch_size=c["batch_size"], iters=c["iters"], weight=c["weight"],
                         b_range=c["b_range"], warmup=c["warmup"], input_prob=c["input_prob"], act_quant=False, plan=plan)
        for n, mm in m.named_modules():
            if isinstance(mm, ow.QuantModule) and mm.org_weight is not None:
                key = f"{full}.{n}" if n else full
                assert torch.equal(mm.weight_quantizer.alpha.data, walk["alpha"][key]), key
                assert torch.equal(mm.weight_quantizer(mm.weight).detach(), walk["hard"][key]), key
                seen += 1
    assert seen == len(walk["alpha"])
    q.eval()
    q.set_quant_state(True, True)
    MG._output_layer(q, walk["is_cheng"]).set_quant_state(True, False)
    with torch.no_grad():
        out = q(x)
    assert torch.equal(out["x_hat"], walk["final"]["x_hat"]) and torch.equal(out["likelihoods"]["y"], walk["final"]["lik_y"])


def test_loss_function_matches_the_reference(golden_w):
    """LossFunction.__call__ (layer_opt.py:114-173): rec + task + annealed rounding regulariser across the warm-up."""
    g = golden_w["loss_function"]
    fp, x = MG.build_fp_model(*MG.WRAP_ARCHS[0][:3])
    q = ow.QuantModel(copy.deepcopy(fp), MG.WQ8, MG.AQ8).eval()
    q.set_quant_state(True, False)
    with torch.no_grad():
        q(x)
    layer = q.model.g_a[2]
    layer.weight_quantizer = oq.AdaRoundQuantizer(layer.weight_quantizer, layer.org_weight.data)
    layer.weight_quantizer.alpha.data += g["alpha_shift"]
    f = ocal.LossFunction(layer, 0.01, 10, (20, 2), 0.2, 2.0, 2.0)
    vals = torch.stack([f(g["pred"], g["tgt"], g["pred"], g["tgt"]).detach() for _ in range(10)])
    assert torch.equal(vals, g["values"])
    assert vals[0] < vals[2]                     # the regulariser switches on at count >= warmup * max_count


def test_lu_quant_model_matches_the_reference(golden_w):
    """LU quant_int.QuantModule / QuantModel / QuantCodingModel (quant_layer.py:79-137, quant_model.py:9-78)."""
    g = golden_w["lu_model"]
    fp, x = MG.build_fp_model(g["arch"], g["kw"], g["gain"])
    q = ow.LUQuantModel(copy.deepcopy(fp), MG.WQ8, g["aq"]).eval()
    q.set_quant_state(True, True)
    q.disable_network_output_quantization()
    with torch.no_grad():
        out = q(x)
    assert torch.equal(out["x_hat"], g["x_hat"]) and torch.equal(out["likelihoods"]["y"], g["lik_y"])
    assert torch.equal(out["likelihoods"]["z"], g["lik_z"])
    mods = {n: m for n, m in q.named_modules() if isinstance(m, ow.LUQuantModule)}
    assert list(mods) == list(g["weights_u8"])
    for n, m in mods.items():
        assert m.weight.dtype == torch.uint8 and torch.equal(m.weight.data, g["weights_u8"][n]), n
    qc = ow.LUQuantModel(copy.deepcopy(fp), MG.WQ8, g["aq"], skip_prefixes=("g_a", "g_s"))
    assert [n for n, m in qc.named_modules() if isinstance(m, ow.LUQuantModule)] == g["coding_modules"]


@pytest.mark.skipif(not _ref_shim.available(), reason="/root/reference is only present in the build container")
def test_shim_imports_the_unmodified_reference_packages():
    TO = _ref_shim.import_task_oriented()
    assert TO.QuantModel.__module__ == "quantization.quant_model"
    assert os.path.realpath(TO.__file__).startswith("/root/reference/")
    LU = _ref_shim.import_light_uniform()
    assert os.path.realpath(LU.__file__).startswith("/root/reference/")
