"""Parity checker: the CUDA path against the oracle on one model / image.  TEST INFRASTRUCTURE ONLY.

Used by tests/test_gpu_reference_parity.py (BASELINE-size parity tests) and by bench.py's `cpu_baseline` leg, which puts
the returned dict on the bench line as `parity`.  The oracle side is the pinned restatement (oracle/__init__.py); the
product side is reached only through `rdo_ptq_b200`'s public classes.

Bars (BASELINE.json north_star): integer weight codes bit-exact; per-layer outputs within 1e-4 relative on the same
inputs; end-to-end bpp within 1e-3 and PSNR within 0.01 dB.
"""
import time

import torch

from . import codec as ocodec, evalpath as oeval, quant_wrap as owrap, quantizers as oq

WQ8 = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ8 = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _rows(model, x, kinds):
    rows, hooks = [], []
    for name, m in model.named_modules():
        if isinstance(m, kinds):
            hooks.append(m.register_forward_hook(
                lambda _m, i, o, name=name: rows.append((name, i[0].detach().clone(), o.detach().clone()))))
    with torch.no_grad():
        res = model(x)
    for h in hooks:
        h.remove()
    return res, rows


def compare_forward(arch, kw, gain, hw, dev, wq=None, aq=None, lu=False, follow_bits=False, engine=None, pad=256,
                    layer_checks=True, self_noise=False):
    """Builds the oracle codec (CPU) and the product codec (CUDA) with identical seeded random-init parameters, wraps both
    (TO rules, or LU rules with `lu=True`), and compares

      * integer weight codes of every wrapped layer                                        -> `codes_equal`
      * W-only forward: every layer fed the ORACLE's input                                 -> `worst_layer_rel_err`
      * W-only and W+A end-to-end bpp / PSNR                                               -> `d_bpp_*`, `d_psnr_*`
      * W+A forward, layer-local: share of activation codes that differ (boundary flips)   -> `a8_flip_rate`
      * `self_noise`: the oracle against itself with the other CPU conv backend            -> `ref_self_noise_*`

    `engine`: "simt" (exact fp32 CUDA engine), "tc" / "auto" (tcgen05 split-bf16 engine) or None (leave as is)."""
    from rdo_ptq_b200 import codec, ops, quantization as Q, quant_int as LU, synth, evaluate as E
    from rdo_ptq_b200.quantization.quantizer import UniformAffineQuantizer as PUAQ
    wq, aq = dict(wq or WQ8), dict(aq or AQ8)
    prev_engine = ops.DEFAULT_ENGINE
    if engine is not None:
        ops.set_default_engine(engine)
    if follow_bits:
        oq.UniformAffineQuantizer.act_bits_follow_n_bits = PUAQ.act_bits_follow_n_bits = True
    try:
        torch.manual_seed(1005)
        om = ocodec.ARCHS[arch](**kw).eval()
        synth.init_weights(om, gain=gain)
        pm = codec.ARCHS[arch](**kw).eval()
        pm.load_state_dict(om.state_dict())
        pm.to(dev)
        x = synth.synthetic_image(*hw)
        xp, xg = oeval.pad(x, pad), E.pad(x.to(dev), pad)
        t0 = time.perf_counter()
        with torch.no_grad():
            om(xp), pm(xg)                                     # FP forward first: bakes the MaskedConv2d mask (Q5)
        if lu:
            oqm, pqm = owrap.LUQuantModel(om, wq, aq).eval(), LU.QuantModel(pm, wq, aq).eval()
            okind, pkind = (owrap.LUQuantModule,), (LU.QuantModule,)
        else:
            oqm, pqm = owrap.QuantModel(om, wq, aq).eval(), Q.QuantModel(pm, wq, aq).eval()
            okind, pkind = (owrap.QuantModule,), (Q.QuantModule,)
        res = dict(arch=arch, kw=kw, hw=list(hw), n_bits_w=wq["n_bits"], rules="LU" if lu else "TO",
                   engine=engine or "default")
        pmods = dict((n, m) for n, m in pqm.named_modules() if isinstance(m, pkind))
        omods = dict((n, m) for n, m in oqm.named_modules() if isinstance(m, okind))
        assert list(pmods) == list(omods), "graph rewrite differs"

        def metrics(o_out, p_out):
            return (E.compute_bpp(p_out) - oeval.compute_bpp(o_out),
                    E.compute_psnr(E.crop(p_out["x_hat"], hw), x.to(dev), clamp=True) -
                    oeval.compute_psnr(x, oeval.crop(o_out["x_hat"], hw).clamp(0, 1)),
                    oeval.compute_bpp(o_out), oeval.compute_psnr(x, oeval.crop(o_out["x_hat"], hw).clamp(0, 1)))

        if lu:                                                 # LU: uint8 codes + Q8.8 from the first forward on
            oqm.set_quant_state(True, True)
            pqm.set_quant_state(True, True)
            oqm.disable_network_output_quantization()
            pqm.disable_network_output_quantization()
            ref, rows = _rows(oqm, xp, okind)
            with torch.no_grad():
                out = pqm(xg)
            res["codes_equal"] = all(torch.equal(pmods[n].weight.data.cpu(), omods[n].weight.data) for n in pmods)
            worst = 0.0
            if layer_checks:
                for name, xi, yo in rows:
                    with torch.no_grad():
                        yi = pmods[name](xi.to(dev)).cpu()
                    # Q8.8 grid: a value may land on the neighbouring grid point when the pre-quant value sits on a boundary
                    worst = max(worst, ((yi - yo).abs() > 1.0 / 256 + 1e-7).float().mean().item())
            res["q88_off_grid_rate"] = worst
            res["d_bpp_wa"], res["d_psnr_wa"], res["bpp_ref"], res["psnr_ref"] = metrics(ref, out)
            res["seconds"] = time.perf_counter() - t0
            return res
        oqm.set_quant_state(True, False)
        pqm.set_quant_state(True, False)
        ref, rows = _rows(oqm, xp, okind)
        with torch.no_grad():
            out = pqm(xg)
        eq = True
        for n, m in pmods.items():
            if m.weight is not None:
                o = omods[n]
                eq &= torch.equal(m.weight_quantizer.delta.cpu().reshape(-1), o.weight_quantizer.delta.reshape(-1))
                eq &= torch.equal(m.weight_quantizer.codes(m.weight).cpu(), o.weight_quantizer.codes(o.weight))
        res["codes_equal"] = bool(eq)
        worst, worst_name = 0.0, ""
        if layer_checks:
            for name, xi, yo in rows:
                with torch.no_grad():
                    e = rel_err(pmods[name](xi.to(dev)), yo)
                if e > worst:
                    worst, worst_name = e, name
        res["worst_layer_rel_err"], res["worst_layer"] = worst, worst_name
        res["d_bpp_w"], res["d_psnr_w"], res["bpp_ref_w"], res["psnr_ref_w"] = metrics(ref, out)
        # W + A: dynamic activation quantisers on for trained units, output layer weights-only (main2.py:272-282)
        last = (lambda q: q.model.g_s[-1][0]) if arch.startswith("cheng") else (lambda q: q.model.g_s[-1])
        for q in (oqm, pqm):
            for m in q.modules():
                if hasattr(m, "trained"):
                    m.trained = True
            q.set_quant_state(True, True)
            last(q).set_quant_state(True, False)
        ref8, rows8 = _rows(oqm, xp, okind)
        with torch.no_grad():
            out8 = pqm(xg)
        flips = tot = 0
        levels = (2 ** aq["n_bits"] - 1) if follow_bits else 255
        if layer_checks:
            for name, xi, yo in rows8:
                with torch.no_grad():
                    yi = pmods[name](xi.to(dev)).cpu()
                step = (yo.amax(dim=(0, 2, 3), keepdim=True) - yo.amin(dim=(0, 2, 3), keepdim=True)) / levels + 1e-12
                d = (yi - yo).abs()
                flips += (d > 0.5 * step).sum().item()
                tot += d.numel()
        res["a8_flip_rate"] = flips / max(tot, 1)
        res["d_bpp_wa"], res["d_psnr_wa"], res["bpp_ref_wa"], res["psnr_ref_wa"] = metrics(ref8, out8)
        if self_noise:
            # The reference against ITSELF: the same PyTorch-CPU arithmetic through the other convolution backend (oneDNN
            # off).  Conv outputs then differ in the last bits (summation order), activation codes flip where a value
            # sits on a rounding boundary, the flips move the dynamic per-channel ranges and the latent rounding, and
            # the end-to-end metrics move: this is the reproducibility floor of the W+A forward on this model.
            with torch.backends.mkldnn.flags(enabled=False), torch.no_grad():
                alt = oqm(xp)
            res["ref_self_noise_bpp"] = abs(oeval.compute_bpp(alt) - res["bpp_ref_wa"])
            res["ref_self_noise_psnr"] = abs(oeval.compute_psnr(x, oeval.crop(alt["x_hat"], hw).clamp(0, 1)) -
                                             res["psnr_ref_wa"])
        yo_, yp_ = ref8["likelihoods"]["y"], out8["likelihoods"]["y"].cpu()
        res["latent_lik_mismatch"] = ((yo_ - yp_).abs() > 1e-3 * yo_.abs() + 1e-6).float().mean().item()
        res["seconds"] = time.perf_counter() - t0
        return res
    finally:
        ops.DEFAULT_ENGINE = prev_engine
        if follow_bits:
            oq.UniformAffineQuantizer.act_bits_follow_n_bits = PUAQ.act_bits_follow_n_bits = False
