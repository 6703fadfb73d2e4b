"""Pins the oracle against the reference and writes the golden fixtures under tests/golden/.

Run HERE (the container that has /root/reference):   python -m oracle.make_golden
  1. imports the reference's own torch-only files
        task-oriented-PTQ/quantization/quantizer.py   and   light-uniform-PTQ/quant_int/quantizer.py
     straight from /root/reference (importlib, nothing is copied) and asserts that `oracle.quantizers` reproduces
     them BIT-EXACTLY on seeded tensors (weight range/codes for conv, tconv, 2-D gamma and 1-D tensors, AdaRound
     init/soft/hard, dynamic activation quant, LU codes and Q8.8, lp_loss);
  2. stores the reference's outputs as `tests/golden/quantizer_ref.pt` (the pinned vectors);
  3. stores oracle-only vectors for the compressai restatement (`codec_oracle.pt`: GDN, EntropyBottleneck,
     GaussianConditional, a tiny end-to-end mbt2018-mean / cheng2020-attn forward) -- UNPINNED: nothing in the
     reference fixes them, they only guard against drift of the restatement.
/root/reference does not exist on the GPU box: tests read the committed .pt files, never this script's inputs.
"""
import importlib.util
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import quantizers as oq, codec, quant_wrap, evalpath   # noqa: E402
from rdo_ptq_b200 import synth                                     # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _same(a, b, what):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f"{what}: oracle differs from the reference (max |d| = {(a - b).abs().max()})"


def quantizer_vectors():
    warnings.filterwarnings("ignore")
    TO = _load(f"{REF}/task-oriented-PTQ/quantization/quantizer.py", "ref_to_quantizer")
    LU = _load(f"{REF}/light-uniform-PTQ/quant_int/quantizer.py", "ref_lu_quantizer")
    g = torch.Generator().manual_seed(1005)
    G = {}
    cases = {
        "conv": (torch.randn(6, 4, 3, 3, generator=g) * 0.3, False),
        "tconv": (torch.randn(4, 6, 5, 5, generator=g) * 0.1, True),
        "gamma": (torch.rand(5, 5, generator=g) * 0.2, False),
        "vec": (torch.randn(17, generator=g), False),
        "conv_pos": (torch.rand(3, 2, 3, 3, generator=g) + 0.5, False),       # min clamps to 0
        "conv_const": (torch.zeros(2, 2, 1, 1), False),                        # delta floors at 1e-8
    }
    for name, (w, tconv) in cases.items():
        for bits in (8, 4):
            for method in ("max", "max_scale"):
                ref = TO.UniformAffineQuantizer(bits, False, True, method, tconv=tconv)
                mine = oq.UniformAffineQuantizer(bits, False, True, method, tconv=tconv)
                r, m = ref(w.clone()), mine(w.clone())
                _same(ref.delta, mine.delta, f"{name}/delta")
                _same(ref.zero_point, mine.zero_point, f"{name}/zp")
                _same(r, m, f"{name}/dequant")
                codes = torch.clamp(torch.round(w / ref.delta) + ref.zero_point, 0, ref.n_levels - 1)
                _same(codes, mine.codes(w), f"{name}/codes")
                key = f"uaq/{name}/b{bits}/{method}"
                G[key] = dict(w=w, tconv=tconv, bits=bits, method=method, delta=ref.delta, zp=ref.zero_point,
                              dequant=r, codes=codes)
                if method == "max":
                    ar = TO.AdaRoundQuantizer(ref, w.clone(), "learned_hard_sigmoid")
                    am = oq.AdaRoundQuantizer(mine, w.clone())
                    _same(ar.alpha.data, am.alpha.data, f"{name}/alpha")
                    alpha0 = ar.alpha.data.clone()
                    shift = torch.randn(w.shape, generator=g) * 2          # move alpha off its init
                    ar.alpha.data += shift
                    am.alpha.data += shift
                    out = {}
                    for soft in (True, False):
                        ar.soft_targets = am.soft_targets = soft
                        a, b = ar(w), am(w)
                        _same(a.detach(), b.detach(), f"{name}/adaround soft={soft}")
                        out["soft" if soft else "hard"] = a.detach()
                    _same(ar.get_soft_targets().detach(), am.get_soft_targets().detach(), f"{name}/h(alpha)")
                    G[key].update(alpha0=alpha0, alpha=ar.alpha.data.clone(), ada_soft=out["soft"],
                                  ada_hard=out["hard"], h=ar.get_soft_targets().detach())
    # per-tensor (channel_wise=False) and symmetric
    for sym in (False, True):
        w = torch.randn(4, 3, 3, 3, generator=g)
        ref, mine = TO.UniformAffineQuantizer(8, sym, False, "max"), oq.UniformAffineQuantizer(8, sym, False, "max")
        r, m = ref(w.clone()), mine(w.clone())
        _same(r, m, f"per-tensor sym={sym}")
        G[f"uaq/per_tensor/sym{int(sym)}"] = dict(w=w, sym=sym, delta=ref.delta, zp=ref.zero_point, dequant=r)
    # dynamic activation quant (4-D, 2-D) incl. a constant channel (range floors at 1e-6)
    x4 = torch.randn(2, 5, 7, 9, generator=g) * torch.tensor([0.1, 1, 5, 30, 0]).view(1, 5, 1, 1) + 0.3
    x2 = torch.randn(6, 4, generator=g)
    for name, x in (("x4", x4), ("x2", x2)):
        r = TO.ActQuantizer(x.clone())
        _same(r, oq.act_quant(x), f"actq/{name} vectorised")
        _same(r, oq.act_quant_loop(x), f"actq/{name} loop")
        G[f"actq/{name}"] = dict(x=x, out=r)
    # LU: codes + Q8.8
    w = torch.randn(6, 4, 3, 3, generator=g) * 0.2
    ref, mine = LU.UniformAffineQuantizer(8, False, True, "max"), oq.LUUniformAffineQuantizer(8, False, True, "max")
    (rc, rd), (mc, md) = ref(w.clone()), mine(w.clone())
    _same(rc, mc, "lu codes")
    _same(rd, md, "lu delta")
    xq = torch.randn(3, 4, 5, 5, generator=g) * 80
    _same(LU.ActQuantizer(xq), oq.lu_act_quantizer(xq), "lu q8.8")
    G["lu"] = dict(w=w, codes=rc, delta=rd, zp=ref.zero_point, x=xq, q88=LU.ActQuantizer(xq))
    # lp_loss, round_ste
    a, b = torch.randn(2, 3, 4, 4, generator=g), torch.randn(2, 3, 4, 4, generator=g)
    for p in (2.0, 1.0, 2.4):
        _same(TO.lp_loss(a, b, p), oq.lp_loss(a, b, p), f"lp_loss p={p}")
        G[f"lp/{p}"] = dict(a=a, b=b, loss=TO.lp_loss(a, b, p))
    _same(TO.round_ste(a * 3), oq.round_ste(a * 3), "round_ste")
    # search- / moment-based ranges (TO quantizer.py:300-370; LU :265-278).  Own generator: the cases above keep their
    # draws.  Gaussian and heavy-tailed slices at 8 / 4 / 3 bits, so that the searches stop at many different candidates.
    g2 = torch.Generator().manual_seed(2005)

    def heavy(*shape):
        return torch.randn(*shape, generator=g2) * 0.1 * torch.exp(torch.randn(*shape, generator=g2))

    scases = {"conv": (torch.randn(12, 8, 5, 5, generator=g2) * 0.1, False), "tconv": (heavy(8, 12, 5, 5), True),
              "gamma": (heavy(9, 9).abs(), False), "vec": (heavy(33), False)}
    picks = set()
    for name, (w, tconv) in scases.items():
        for bits in (8, 4, 3):
            for method, sym in (("mse", False), ("l1", False), ("l2", False), ("gaussian", False), ("gaussian", True)):
                ref = TO.UniformAffineQuantizer(bits, sym, True, method, tconv=tconv)
                mine = oq.UniformAffineQuantizer(bits, sym, True, method, tconv=tconv)
                r, m = ref(w.clone()), mine(w.clone())
                _same(ref.delta, mine.delta, f"{name}/{method}/delta")
                _same(ref.zero_point, mine.zero_point, f"{name}/{method}/zp")
                _same(r, m, f"{name}/{method}/dequant")
                G[f"uaq_search/{name}/b{bits}/{method}/sym{int(sym)}"] = dict(
                    w=w, tconv=tconv, bits=bits, method=method, sym=sym, delta=ref.delta, zp=ref.zero_point, dequant=r)
                if method in ("mse", "l1", "l2") and w.dim() > 1:
                    full = (w.amax(dim=(0, 2, 3) if tconv else tuple(range(1, w.dim()))) -
                            w.amin(dim=(0, 2, 3) if tconv else tuple(range(1, w.dim())))) / (2 ** bits - 1)
                    picks.update(torch.round((1 - ref.delta.flatten() / full) / 0.05).int().tolist())
    assert len(picks) > 5, f"the shrink search always picked the same candidates {picks}: weak vectors"
    w = heavy(6, 4, 3, 3)
    ref, mine = LU.UniformAffineQuantizer(8, False, True, "mse"), oq.LUUniformAffineQuantizer(8, False, True, "mse")
    (rc, rd), (mc, md) = ref(w.clone()), mine(w.clone())
    _same(rc, mc, "lu mse codes")
    _same(rd, md, "lu mse delta")
    _same(ref.zero_point, mine.zero_point, "lu mse zp")
    G["lu_mse"] = dict(w=w, codes=rc, delta=rd, zp=ref.zero_point)
    return G


def codec_vectors():
    torch.manual_seed(1005)
    G = {}
    gdn = codec.GDN(6)
    with torch.no_grad():
        gdn.gamma.add_(torch.rand(6, 6) * 0.05)
    x = torch.randn(2, 6, 5, 7)
    G["gdn"] = dict(state=gdn.state_dict(), x=x, y=gdn(x).detach())
    igdn = codec.GDN(6, inverse=True)
    igdn.load_state_dict(gdn.state_dict())
    G["igdn"] = dict(y=igdn(x).detach())
    eb = codec.EntropyBottleneck(5).eval()
    with torch.no_grad():
        for i in range(4):
            getattr(eb, f"_factor{i}").uniform_(-0.5, 0.5)
        eb.quantiles[:, 0, 1].uniform_(-1, 1)
    z = torch.randn(2, 5, 4, 6) * 3
    zh, zl = eb(z)
    G["eb"] = dict(state=eb.state_dict(), z=z, z_hat=zh.detach(), lik=zl.detach())
    gc = codec.GaussianConditional(None).eval()
    y, sc, mu = torch.randn(2, 4, 6, 6) * 4, torch.rand(2, 4, 6, 6) * 3, torch.randn(2, 4, 6, 6)
    yh, yl = gc(y, sc, means=mu)
    yh0, yl0 = gc(y, sc)
    G["gc"] = dict(y=y, scales=sc, means=mu, y_hat=yh, lik=yl, y_hat0=yh0, lik0=yl0)
    for arch, kw, gain in (("mbt2018-mean", dict(N=8, M=12), 1.2), ("bmshj2018-hyperprior", dict(N=8, M=12), 1.2),
                           ("cheng2020-attn", dict(N=12), 0.6)):
        m = codec.ARCHS[arch](**kw).eval()
        synth.init_weights(m, gain=gain)
        x = synth.synthetic_image(64, 64)
        with torch.no_grad():
            out = m(x)
        G[f"model/{arch}"] = dict(kw=kw, state=m.state_dict(), x=x, x_hat=out["x_hat"], lik_y=out["likelihoods"]["y"],
                                  lik_z=out["likelihoods"]["z"], bpp=evalpath.compute_bpp(out),
                                  psnr=evalpath.compute_psnr(x, out["x_hat"].clamp(0, 1)))
    return G


# ------------------------------------------------------------------------------------------------------------------
# wrapper / calibration vectors: the reference's OWN quantization + quant_int packages, imported through _ref_shim
WQ8 = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ8 = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
WRAP_ARCHS = (("mbt2018-mean", dict(N=8, M=12), 1.2, False), ("bmshj2018-hyperprior", dict(N=8, M=12), 1.2, False),
              ("cheng2020-attn", dict(N=12), 0.6, True))
CALIB = dict(n_samples=6, patch=64, batch_size=2, iters=24, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)


def build_fp_model(arch, kw, gain):
    """The seeded random-init codec + test image every wrapper vector starts from (one FP forward bakes the masks, Q5)."""
    torch.manual_seed(1005)
    m = codec.ARCHS[arch](**kw).eval()
    synth.init_weights(m, gain=gain)
    x = synth.synthetic_image(64, 64)
    with torch.no_grad():
        m(x)
    return m, x


def _mark_trained(q, kinds):
    for m in q.modules():
        if isinstance(m, kinds):
            m.trained = True


def _output_layer(q, is_cheng):
    return q.model.g_s[-1][0] if is_cheng else q.model.g_s[-1]          # main2.py:258-263


def _layer_outputs(q, x, kinds):
    """{module path: output} of every wrapped module / block of one forward (hooks; order = execution order)."""
    outs, hooks = {}, []
    for name, m in q.named_modules():
        if isinstance(m, kinds):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, name=name: outs.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        res = q(x)
    for h in hooks:
        h.remove()
    return res, outs


def _structure(q, QM, QB):
    return [(n, type(m).__name__, type(m.activation_function).__name__, bool(getattr(m, "disable_act_quant", False)))
            for n, m in q.named_modules() if isinstance(m, (QM, QB))]


def wrap_vectors():
    """Runs the reference's own QuantModel / QuantModule / blocks / save_inp_oup_data / LossFunction /
    layer_reconstruction / block_reconstruction (CPU, via oracle/_ref_shim.py), asserts that oracle.quant_wrap and
    oracle.calib reproduce them BIT-EXACTLY, and returns the REFERENCE's outputs."""
    import copy
    import contextlib
    import io
    import types
    from oracle import _ref_shim as S, calib as ocal
    warnings.filterwarnings("ignore")
    TO = S.import_task_oriented()
    from quantization import quant_layer as r_ql, quant_block as r_qb, utils as r_ut, layer_opt as r_lo, block_opt as r_bo
    G = {}
    # ---- (1) QuantModel forwards: FP, W8, W8A8, W4A8 with 8-bit head/stem, per wrapped module --------------------
    for arch, kw, gain, is_cheng in WRAP_ARCHS:
        fp, x = build_fp_model(arch, kw, gain)
        for tag, wq, head8 in (("w8", WQ8, False), ("w4", dict(WQ8, n_bits=4), True)):
            r = TO.QuantModel(copy.deepcopy(fp), wq, AQ8, is_cheng=is_cheng).eval()
            o = quant_wrap.QuantModel(copy.deepcopy(fp), wq, AQ8, is_cheng=is_cheng).eval()
            rk, ok_ = (r_ql.QuantModule, r_qb.BaseQuantBlock), (quant_wrap.QuantModule, quant_wrap.BaseQuantBlock)
            assert _structure(r, *rk) == _structure(o, *ok_), f"{arch}: graph rewrite differs"
            assert list(r.state_dict().keys()) == list(o.state_dict().keys()), f"{arch}: state-dict keys differ"
            if head8:
                r.set_first_last_layer_to_8bit()
                o.set_first_last_layer_to_8bit()
            r.disable_network_output_quantization()
            o.disable_network_output_quantization()
            case = dict(arch=arch, kw=kw, gain=gain, is_cheng=is_cheng, wq=wq, head8=head8, x=x,
                        structure=_structure(r, *rk))
            for state, trained in (("fp", False), ("w", False), ("wa", True)):
                wqs, aqs = state != "fp", state == "wa"
                for q, kinds in ((r, rk), (o, ok_)):
                    if trained:
                        _mark_trained(q, kinds)
                    q.set_quant_state(wqs, aqs)
                    if aqs:
                        _output_layer(q, is_cheng).set_quant_state(True, False)
                (ro, rl), (oo, ol) = _layer_outputs(r, x, rk), _layer_outputs(o, x, ok_)
                _same(ro["x_hat"], oo["x_hat"], f"{arch}/{tag}/{state}/x_hat")
                for k in ("y", "z"):
                    _same(ro["likelihoods"][k], oo["likelihoods"][k], f"{arch}/{tag}/{state}/lik_{k}")
                assert list(rl) == list(ol), f"{arch}: execution order differs"
                for k in rl:
                    _same(rl[k], ol[k], f"{arch}/{tag}/{state}/{k}")
                # per-module outputs are kept where the GPU tests read them (fixture size): W8 and W8A8 of the 8-bit case
                keep_layers = tag == "w8" and (state == "wa" or (state == "w" and not is_cheng))
                case[state] = dict(x_hat=ro["x_hat"], lik_y=ro["likelihoods"]["y"], lik_z=ro["likelihoods"]["z"],
                                   layers=rl if keep_layers else {}, bpp=evalpath.compute_bpp(ro),
                                   psnr=evalpath.compute_psnr(x, ro["x_hat"].clamp(0, 1)))
            case["n_bits"] = [(n, m.weight_quantizer.n_bits, m.act_quantizer.n_bits, m.disable_act_quant)
                              for n, m in r.named_modules() if isinstance(m, r_ql.QuantModule)]
            assert case["n_bits"] == [(n, m.weight_quantizer.n_bits, m.act_quantizer.n_bits, m.disable_act_quant)
                                      for n, m in o.named_modules() if isinstance(m, quant_wrap.QuantModule)]
            case["codes"] = {n: torch.clamp(torch.round(m.weight / m.weight_quantizer.delta) + m.weight_quantizer.zero_point,
                                            0, m.weight_quantizer.n_levels - 1).detach()
                             for n, m in r.named_modules() if isinstance(m, r_ql.QuantModule) and m.weight is not None}
            G[f"model/{arch}/{tag}"] = case
    # ---- (2) calibration pieces and the whole walk (mbt2018-mean layers; cheng2020-attn blocks + layers) ----------
    args = types.SimpleNamespace(lmbda=0.01, task_loss=2.0, arch="Minnen2018")
    cali = synth.calibration_patches(CALIB["n_samples"], CALIB["patch"])
    plan = ocal.DrawPlan(1005)
    kw_ref = dict(cali_data=cali, batch_size=CALIB["batch_size"], iters=CALIB["iters"], weight=CALIB["weight"],
                  input_prob=CALIB["input_prob"], lr=4e-5, asym=True, b_range=CALIB["b_range"], warmup=CALIB["warmup"],
                  act_quant=False, opt_mode="mse", config=None, args=args)
    kw_ora = dict(batch_size=CALIB["batch_size"], iters=CALIB["iters"], weight=CALIB["weight"], b_range=CALIB["b_range"],
                  warmup=CALIB["warmup"], input_prob=CALIB["input_prob"], act_quant=False, plan=plan)
    for arch, kw, gain, is_cheng in (WRAP_ARCHS[0], WRAP_ARCHS[2]):
        fp, x = build_fp_model(arch, kw, gain)
        wq = dict(WQ8, n_bits=4)
        r = TO.QuantModel(copy.deepcopy(fp), wq, AQ8, is_cheng=is_cheng).eval()
        o = quant_wrap.QuantModel(copy.deepcopy(fp), wq, AQ8, is_cheng=is_cheng).eval()
        for q in (r, o):
            q.set_first_last_layer_to_8bit()
            q.disable_network_output_quantization()
            q.set_quant_state(True, False)
            with torch.no_grad():
                q(cali[:CALIB["batch_size"]])                                   # main2.py:197-201 (scale init)
            _output_layer(q, is_cheng).set_quant_state(True, False)
        args.arch = "Cheng2020" if is_cheng else "Minnen2018"
        units_r, units_o = [], []

        def collect(module, prefix, QM, QB, out):                               # main2.py:227-250 walk order
            for name, m in module.named_children():
                full = f"{prefix}.{name}" if prefix else name
                if isinstance(m, (QM, QB)):
                    out.append((full, name, m))
                else:
                    collect(m, full, QM, QB, out)
        collect(r, "", r_ql.QuantModule, r_qb.BaseQuantBlock, units_r)
        collect(o, "", quant_wrap.QuantModule, quant_wrap.BaseQuantBlock, units_o)
        assert [u[0] for u in units_r] == [u[0] for u in units_o]
        # Q1: with compressai's "0", "1", ... child names the coder-name tests never match
        ml, nl = r_lo.find_unquantized_module(r, units_r[1][1], [], [])
        assert ml == [] and nl == [], "find_unquantized_module matched something: SURVEY Q1 no longer holds"
        walk = dict(arch=arch, kw=kw, gain=gain, is_cheng=is_cheng, wq=wq, cali_sum=cali.double().sum(), calib=dict(CALIB), seed=plan.seed,
                    units=[u[0] for u in units_r], find_unquantized=(len(ml), len(nl)), alpha={}, hard={}, caches={},
                    loss={})
        limit = None if not is_cheng else 14            # cheng2020: the g_a units (RBWS, RB, attention convs) + first of h_a
        walk["limit"] = limit
        for uid, ((full, name, mr), (_, _, mo)) in enumerate(zip(units_r, units_o)):
            if limit is not None and uid >= limit:
                break
            if mr.ignore_reconstruction:
                continue
            is_block = isinstance(mr, r_qb.BaseQuantBlock)
            # save_inp_oup_data on its own (utils.py:92-139), before the unit is trained
            with contextlib.redirect_stdout(io.StringIO()):
                (rq, rf), rout = r_ut.save_inp_oup_data(r, mr, cali, True, False, batch_size=1, input_prob=True)
            (oq_, of), oout = ocal.save_inp_oup_data(o, mo, cali, False, is_block)
            _same(rq, oq_, f"{arch}/{full}/quant_in")
            _same(rf, of, f"{arch}/{full}/fp_in")
            _same(rout, oout, f"{arch}/{full}/fp_out")
            if uid == 2:
                walk["caches"][full] = dict(quant_in=rq, fp_in=rf, fp_out=rout)
            if getattr(mr, "org_weight", 1) is None and not is_block:
                with S.cpu_device():
                    r_lo.layer_reconstruction(r, mr, name, **kw_ref)             # PixelShuffle wrapper: returns early
                ocal.reconstruct(o, mo, uid, name, cali, **kw_ora)
                continue
            fn = r_bo.block_reconstruction if is_block else r_lo.layer_reconstruction
            with S.cpu_device(), S.replay_draws(plan, uid, cali.size(0), CALIB["batch_size"], CALIB["input_prob"]), \
                    contextlib.redirect_stdout(io.StringIO()):
                fn(r, mr, name, **kw_ref)
            ocal.reconstruct(o, mo, uid, name, cali, **kw_ora)
            rmods = [m for m in mr.modules() if isinstance(m, r_ql.QuantModule) and m.org_weight is not None]
            omods = [m for m in mo.modules() if isinstance(m, quant_wrap.QuantModule) and m.org_weight is not None]
            names = [n for n, m in mr.named_modules() if isinstance(m, r_ql.QuantModule) and m.org_weight is not None]
            for n, a, b in zip(names, rmods, omods):
                key = f"{full}.{n}" if n else full
                _same(a.weight_quantizer.alpha.data, b.weight_quantizer.alpha.data, f"{arch}/{key}/alpha")
                _same(a.weight_quantizer(a.weight).detach(), b.weight_quantizer(b.weight).detach(), f"{arch}/{key}/hard")
                assert a.trained and b.trained and not a.weight_quantizer.soft_targets
                walk["alpha"][key] = a.weight_quantizer.alpha.data.clone()
                walk["hard"][key] = a.weight_quantizer(a.weight).detach().clone()
        for q, kinds in ((r, (r_ql.QuantModule, r_qb.BaseQuantBlock)), (o, (quant_wrap.QuantModule, quant_wrap.BaseQuantBlock))):
            q.eval()
            q.set_quant_state(True, True)
            _output_layer(q, is_cheng).set_quant_state(True, False)
        with torch.no_grad():
            ro, oo = r(x), o(x)
        _same(ro["x_hat"], oo["x_hat"], f"{arch}/after-walk x_hat")
        walk["final"] = dict(x=x, x_hat=ro["x_hat"], lik_y=ro["likelihoods"]["y"], lik_z=ro["likelihoods"]["z"],
                             bpp=evalpath.compute_bpp(ro), psnr=evalpath.compute_psnr(x, ro["x_hat"].clamp(0, 1)))
        G[f"walk/{arch}"] = walk
    # ---- (3) LossFunction.__call__ on its own (layer_opt.py:114-173), across the warm-up boundary ------------------
    fp, x = build_fp_model(*WRAP_ARCHS[0][:3])
    r = TO.QuantModel(copy.deepcopy(fp), WQ8, AQ8).eval()
    o = quant_wrap.QuantModel(copy.deepcopy(fp), WQ8, AQ8).eval()
    for q in (r, o):
        q.set_quant_state(True, False)
        with torch.no_grad():
            q(x)
    lr_, lo_ = r.model.g_a[2], o.model.g_a[2]
    lr_.weight_quantizer = TO.quantizer.AdaRoundQuantizer(uaq=lr_.weight_quantizer, round_mode="learned_hard_sigmoid",
                                                          weight_tensor=lr_.org_weight.data)
    lo_.weight_quantizer = oq.AdaRoundQuantizer(lo_.weight_quantizer, lo_.org_weight.data)
    g = torch.Generator().manual_seed(77)
    shift = torch.randn(lr_.weight.shape, generator=g)
    lr_.weight_quantizer.alpha.data += shift
    lo_.weight_quantizer.alpha.data += shift
    pred, tgt = torch.randn(2, 8, 6, 6, generator=g), torch.randn(2, 8, 6, 6, generator=g)
    args.arch = "Minnen2018"
    rf_ = r_lo.LossFunction(lr_, round_loss="relaxation", weight=0.01, max_count=10, rec_loss="mse", b_range=(20, 2),
                            decay_start=0, warmup=0.2, p=2.0, lmbda=args.lmbda, metric=args.task_loss)
    of_ = ocal.LossFunction(lo_, 0.01, 10, (20, 2), 0.2, 2.0, 2.0)
    vals = []
    for _ in range(10):
        a, b = rf_(pred, tgt, quant_net_out=pred, cali_data=tgt), of_(pred, tgt, pred, tgt)
        _same(a.detach(), b.detach(), "LossFunction")
        vals.append(a.detach())
    G["loss_function"] = dict(alpha_shift=shift, pred=pred, tgt=tgt, values=torch.stack(vals))
    # ---- (4) light-uniform-PTQ: quant_int.QuantModule / QuantModel on Balle2018 (LU wrapper rules) -------------------
    LU = S.import_light_uniform()
    from quant_int import quant_layer as l_ql
    fp, x = build_fp_model(*WRAP_ARCHS[1][:3])
    aq_lu = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=True)
    r = LU.QuantModel(copy.deepcopy(fp), WQ8, aq_lu).eval()
    o = quant_wrap.LUQuantModel(copy.deepcopy(fp), WQ8, aq_lu).eval()
    for q in (r, o):
        q.set_quant_state(True, True)
        q.disable_network_output_quantization()
    # quant_layer.py:118 assigns uint8 codes to `weight.data`; torch >= 2 only allows that on a parameter that does not
    # require grad (the reference predates the check), so the call site freezes the parameters first -- values unchanged
    for p_ in r.parameters():
        p_.requires_grad_(False)
    with torch.no_grad():
        ro, oo = r.model(x), o.model(x)                  # LU's 2-argument forward is TinyLIC's (Q8): call the codec
    _same(ro["x_hat"], oo["x_hat"], "lu/x_hat")
    _same(ro["likelihoods"]["y"], oo["likelihoods"]["y"], "lu/lik_y")
    lu_w = {n: m.weight.data.clone() for n, m in r.named_modules() if isinstance(m, l_ql.QuantModule)}
    for (n, a), (_, b) in zip(lu_w.items(), ((n, m.weight.data) for n, m in o.named_modules()
                                             if isinstance(m, quant_wrap.LUQuantModule))):
        assert a.dtype == torch.uint8 and torch.equal(a, b), f"lu/{n}: uint8 codes differ"
    rc = LU.QuantCodingModel(copy.deepcopy(fp), WQ8, aq_lu)
    oc = quant_wrap.LUQuantModel(copy.deepcopy(fp), WQ8, aq_lu, skip_prefixes=("g_a", "g_s"))
    lu_coding = [n for n, m in rc.named_modules() if isinstance(m, l_ql.QuantModule)]
    assert lu_coding == [n for n, m in oc.named_modules() if isinstance(m, quant_wrap.LUQuantModule)]
    G["lu_model"] = dict(arch=WRAP_ARCHS[1][0], kw=WRAP_ARCHS[1][1], gain=WRAP_ARCHS[1][2], x=x, aq=aq_lu,
                         x_hat=ro["x_hat"], lik_y=ro["likelihoods"]["y"], lik_z=ro["likelihoods"]["z"], weights_u8=lu_w,
                         coding_modules=lu_coding, bpp=evalpath.compute_bpp(ro),
                         psnr=evalpath.compute_psnr(x, ro["x_hat"].clamp(0, 1)))
    return G


def main():
    os.makedirs(OUT, exist_ok=True)
    q = quantizer_vectors()
    torch.save(q, os.path.join(OUT, "quantizer_ref.pt"))
    print(f"quantizer_ref.pt: {len(q)} cases pinned bit-exactly against /root/reference")
    c = codec_vectors()
    torch.save(c, os.path.join(OUT, "codec_oracle.pt"))
    print(f"codec_oracle.pt: {len(c)} cases (oracle-only, unpinned)")
    w = wrap_vectors()
    torch.save(w, os.path.join(OUT, "wrap_ref.pt"))
    print(f"wrap_ref.pt: {len(w)} groups pinned bit-exactly against the reference's quantization / quant_int packages")


if __name__ == "__main__":
    main()
