"""GPU parity of the drop-in surface (QuantModel / QuantModule / blocks / reconstruction / evaluation) against the
CPU oracle on identical seeded weights and synthetic inputs.

Bars (BASELINE.json north_star): integer weight codes bit-exact; per-layer outputs <= 1e-4 relative; end-to-end bpp
within 1e-3 and PSNR within 0.01 dB."""
import math

import pytest
import torch

from oracle import codec as ocodec, quant_wrap as owrap, calib as ocalib, evalpath as oeval, quantizers as oq
from rdo_ptq_b200 import synth

pytestmark = pytest.mark.gpu

WQ = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_pair(arch, kw, gain, dev, wq=WQ, aq=AQ, state=None):
    """Oracle model on CPU + product model on the GPU with identical parameters, both wrapped in QuantModel after
    one FP forward (bakes the MaskedConv2d mask)."""
    from rdo_ptq_b200 import codec, quantization as Q
    torch.manual_seed(1005)
    om = ocodec.ARCHS[arch](**kw).eval()
    if state is None:
        synth.init_weights(om, gain=gain)
    else:
        om.load_state_dict(state)
    pm = codec.ARCHS[arch](**kw).eval()
    pm.load_state_dict(om.state_dict())
    pm.to(dev)
    return om, pm, Q


def per_layer_io(model, x, kinds):
    """[(name, input, output)] of every module of `kinds`, cloned (the blocks apply in-place LeakyReLU afterwards)."""
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


def layer_local_parity(ref_rows, pmods, dev, check):
    """Feed the ORACLE's input of every layer to the matching CUDA layer (same inputs => no cascade of rounding flips)."""
    for name, x, y in ref_rows:
        with torch.no_grad():
            out = pmods[name](x.to(dev))
        check(name, out.cpu(), y)


@pytest.mark.parametrize("arch,kw,gain", [("mbt2018-mean", dict(N=8, M=12), 1.2),
                                          ("bmshj2018-hyperprior", dict(N=8, M=12), 1.2),
                                          ("cheng2020-attn", dict(N=12), 0.6)])
def test_fp_forward_matches_golden(dev, golden_c, arch, kw, gain):
    g = golden_c[f"model/{arch}"]
    om, pm, _ = build_pair(arch, g["kw"], gain, dev, state=g["state"])
    with torch.no_grad():
        out = pm(g["x"].to(dev))
    assert rel_err(out["x_hat"], g["x_hat"]) < 1e-4
    from rdo_ptq_b200 import evaluate as E
    assert abs(E.compute_bpp(out) - g["bpp"]) < 1e-3
    assert abs(E.compute_psnr(out["x_hat"], g["x"].to(dev), clamp=True) - g["psnr"]) < 0.01


@pytest.mark.parametrize("arch,kw,gain,hw", [("mbt2018-mean", dict(N=32, M=48), 1.2, (128, 192)),
                                             ("cheng2020-attn", dict(N=24), 0.6, (64, 128))])
def test_quantized_forward_per_layer_parity(dev, arch, kw, gain, hw):
    from rdo_ptq_b200 import evaluate as E
    om, pm, Q = build_pair(arch, kw, gain, dev)
    x = synth.synthetic_image(*hw)
    with torch.no_grad():
        om(x), pm(x.to(dev))                                   # FP forward first (Q5)
    oq_model, pq_model = owrap.QuantModel(om, WQ, AQ).eval(), Q.QuantModel(pm, WQ, AQ).eval()
    oq_model.set_quant_state(True, False)
    pq_model.set_quant_state(True, False)
    ref, ref_rows = per_layer_io(oq_model, x, (owrap.QuantModule,))
    with torch.no_grad():
        out = pq_model(x.to(dev))
    pmods = dict((n, m) for n, m in pq_model.named_modules() if isinstance(m, Q.QuantModule))
    omods = dict((n, m) for n, m in oq_model.named_modules() if isinstance(m, owrap.QuantModule))
    assert list(pmods) == list(omods) and len(ref_rows) > 15
    # integer weight codes: bit-exact for every wrapped layer
    for n, m in pmods.items():
        if m.weight is not None:
            o = omods[n]
            assert torch.equal(m.weight_quantizer.delta.cpu().reshape(-1), o.weight_quantizer.delta.reshape(-1)), n
            assert torch.equal(m.weight_quantizer.codes(m.weight).cpu(), o.weight_quantizer.codes(o.weight)), n
    # W8 (weights only): per-layer outputs within 1e-4 relative on identical inputs
    worst = [0.0]

    def chk(name, a, b):
        worst[0] = max(worst[0], rel_err(a, b))
        assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))
    layer_local_parity(ref_rows, pmods, dev, chk)
    assert abs(E.compute_bpp(out) - oeval.compute_bpp(ref)) < 1e-3
    assert abs(E.compute_psnr(out["x_hat"], x.to(dev), clamp=True) - oeval.compute_psnr(x, ref["x_hat"].clamp(0, 1))) < 0.01
    # W8A8: dynamic activation quant switched on for trained layers (main2.py:272-282)
    for q in (oq_model, pq_model):
        for m in q.modules():
            if hasattr(m, "trained"):
                m.trained = True
        q.set_quant_state(True, True)
    ref8, ref8_rows = per_layer_io(oq_model, x, (owrap.QuantModule,))
    with torch.no_grad():
        out8 = pq_model(x.to(dev))
    flips = [0, 0]

    def chk8(name, a, b):
        # same inputs: activation codes may flip only where the pre-quant value sits on a rounding boundary
        step = (b.amax() - b.amin()).item() / 255 + 1e-12
        d = (a - b).abs()
        assert (d <= 1.01 * step).all(), name
        flips[0] += (d > 0.5 * step).sum().item()
        flips[1] += d.numel()
    layer_local_parity(ref8_rows, pmods, dev, chk8)
    assert flips[0] / flips[1] < 1e-4
    # end to end, rounding flips cascade through the dynamic ranges and the latent rounding: compare the metrics
    bpp_ref, bpp = oeval.compute_bpp(ref8), E.compute_bpp(out8)
    p_ref, p = oeval.compute_psnr(x, ref8["x_hat"].clamp(0, 1)), E.compute_psnr(out8["x_hat"], x.to(dev), clamp=True)
    print(f"W8A8 {arch}: bpp {bpp:.5f} vs {bpp_ref:.5f}; psnr {p:.4f} vs {p_ref:.4f}; worst W8 layer rel err {worst[0]:.2e}")
    assert abs(bpp - bpp_ref) < 0.01 * bpp_ref + 1e-3 and abs(p - p_ref) < 0.1


def test_evaluate_matches_oracle_end_to_end(dev):
    """Test_kodak path (pad-256, crop, clamp, PSNR, bpp) on Kodak-shaped synthetic images, W8 weights."""
    from rdo_ptq_b200 import evaluate as E
    om, pm, Q = build_pair("mbt2018-mean", dict(N=16, M=24), 1.2, dev)
    imgs = synth.synthetic_images(2, 200, 300)
    with torch.no_grad():
        om(oeval.pad(imgs[0], 256)), pm(E.pad(imgs[0].to(dev), 256))
    oqm, pqm = owrap.QuantModel(om, WQ, AQ).eval(), Q.QuantModel(pm, WQ, AQ).eval()
    oqm.set_quant_state(True, False)
    pqm.set_quant_state(True, False)
    ps, bs = oeval.evaluate(oqm, imgs)
    res = E.evaluate(pqm, [i.to(dev) for i in imgs])
    assert res["count"] == 2
    assert abs(res["psnr"] - sum(ps) / 2) < 0.01 and abs(res["bpp"] - sum(bs) / 2) < 1e-3
    rd = E.RateDistortionLoss(lmbda=0.01)(pqm(E.pad(imgs[0].to(dev), 256)), E.pad(imgs[0].to(dev), 256))
    with torch.no_grad():
        ref_rd = oeval.rate_distortion_loss(oqm(oeval.pad(imgs[0], 256)), oeval.pad(imgs[0], 256), 0.01)
    assert abs(rd["loss"] - float(ref_rd["loss"])) < 1e-3 * max(1.0, float(ref_rd["loss"]))


def test_lu_uint8_q88_forward_parity(dev):
    """BASELINE config 1: light-uniform-PTQ rules on Balle2018 (uint8 per-channel weights, Q8.8 activations, GDN fp32)."""
    from rdo_ptq_b200 import quant_int as LU, evaluate as E
    om, pm, _ = build_pair("bmshj2018-hyperprior", dict(N=16, M=24), 1.2, dev)
    x = synth.synthetic_image(128, 192)
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=True)
    oqm, pqm = owrap.LUQuantModel(om, wq, aq).eval(), LU.QuantModel(pm, wq, aq).eval()
    oqm.set_quant_state(True, True)
    pqm.set_quant_state(True, True)
    ref, ref_rows = per_layer_io(oqm, x, (owrap.LUQuantModule,))
    with torch.no_grad():
        out = pqm(x.to(dev))
    om_mods = [m for m in oqm.modules() if isinstance(m, owrap.LUQuantModule)]
    pm_mods = [m for m in pqm.modules() if isinstance(m, LU.QuantModule)]
    assert len(om_mods) == len(pm_mods) == 14
    for a, b in zip(pm_mods, om_mods):
        assert a.weight.dtype == torch.uint8 and torch.equal(a.weight.data.cpu(), b.weight.data)      # bit-exact codes
    pmods = dict((n, m) for n, m in pqm.named_modules() if isinstance(m, LU.QuantModule))
    flips = [0, 0]

    def chk(name, a, b):
        d = (a - b).abs()
        assert (d <= 1.0 / 256 + 1e-6).all(), name                # same inputs: at most one Q8.8 step, on boundaries
        flips[0] += (d > 0).sum().item()
        flips[1] += d.numel()
    layer_local_parity(ref_rows, pmods, dev, chk)
    assert flips[0] / flips[1] < 1e-3
    bpp_ref, bpp = oeval.compute_bpp(ref), E.compute_bpp(out)
    print(f"LU Q8.8: bpp {bpp:.5f} vs {bpp_ref:.5f}; boundary flips {flips[0]}/{flips[1]}")
    assert abs(bpp - bpp_ref) < 0.01 * bpp_ref + 1e-3


class ReplayPlan:
    """Feeds the oracle's CPU-drawn (idx, mask) pairs to the CUDA loop."""

    def __init__(self, seed=1005):
        self.cpu = ocalib.DrawPlan(seed)

    def draw(self, unit_id, it, n, bs, shape, prob, device):
        idx, keep = self.cpu.draw(unit_id, it, n, bs, shape, prob)
        return idx.to(device), (None if keep is None else keep.to(device)), 0


def _calib_pair(dev, arch, kw, gain, bits=4):
    wq = dict(n_bits=bits, channel_wise=True, scale_method="max")
    om, pm, Q = build_pair(arch, kw, gain, dev)
    cali = synth.calibration_patches(4, 64)
    with torch.no_grad():
        om(cali[:1]), pm(cali[:1].to(dev))
    oqm, pqm = owrap.QuantModel(om, wq, AQ).eval(), Q.QuantModel(pm, wq, AQ).eval()
    for q, c in ((oqm, cali), (pqm, cali.to(dev))):
        q.set_quant_state(True, False)
        with torch.no_grad():
            q(c[:2])                                            # scale init (main2.py:194-198)
    return oqm, pqm, Q, cali


@pytest.mark.parametrize("layer_path", ["g_a.2", "g_a.1", "g_s.2", "h_s.0"])
def test_layer_reconstruction_parity(dev, layer_path):
    """First-iteration dL/dalpha and the alpha trajectory over 30 iterations of one AdaRound problem."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
    sub, idx = layer_path.split(".")
    olayer, player = getattr(oqm.model, sub)[int(idx)], getattr(pqm.model, sub)[int(idx)]
    # the units before this one count as already reconstructed, so quant_in != fp_in and the loss is not fp noise
    for qm in (oqm, pqm):
        for j in range(int(idx)):
            m = getattr(qm.model, sub)[j]
            if hasattr(m, "trained"):
                m.trained = True
    kw = dict(batch_size=2, iters=30, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)
    otrace, ptrace = {}, {}
    ocalib.reconstruct(oqm, olayer, 3, idx, cali, plan=ocalib.DrawPlan(), trace=otrace, **kw)

    class Args:
        task_loss = 2.0
    Q.layer_reconstruction(pqm, player, idx, cali.to(dev), asym=True, act_quant=False, opt_mode='mse', args=Args(),
                           plan=ReplayPlan(), unit_id=3, trace=ptrace, **kw)
    assert rel_err(ptrace["out"], otrace["out0"]) < 1e-4
    # elements whose soft target starts exactly on the clamp boundary (rest == 0: e.g. each row's extreme weights) have a
    # gradient gate that is a float coin-flip in the reference itself; parity is defined on the interior elements
    h0 = otrace["h0"]
    inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
    assert inner.float().mean() > 0.8
    assert rel_err(ptrace["d_alpha"][0].cpu()[inner], otrace["grad0"][0][inner]) < 1e-3
    a_ref, a_gpu = olayer.weight_quantizer.alpha.data, player.weight_quantizer.alpha.data.cpu()
    assert (a_ref - a_gpu)[inner].abs().max().item() < 5e-3        # 30 Adam steps of 1e-3 each: well inside 1 step
    flips = ((a_ref >= 0) != (a_gpu >= 0))[inner].float().mean().item()
    assert flips < 1e-3
    assert player.trained and not player.weight_quantizer.soft_targets
    # hardened weights: integer codes agree wherever the rounding decision agrees
    oc, pc = olayer.weight_quantizer.codes(olayer.weight), player.weight_quantizer.codes(player.weight).cpu()
    assert ((oc != pc)[inner].float().mean().item()) <= flips + 1e-9


def test_block_reconstruction_parity(dev):
    """Cheng2020 residual blocks: joint AdaRound over all QuantModules of a block (dgrad through conv, GDN, subpel)."""
    oqm, pqm, Q, cali = _calib_pair(dev, "cheng2020-attn", dict(N=12), 0.6)
    kw = dict(batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)

    class Args:
        task_loss = 2.0
    for unit_id, (sub, idx) in enumerate((("g_a", 0), ("g_a", 1), ("g_s", 2))):
        oblk, pblk = getattr(oqm.model, sub)[idx], getattr(pqm.model, sub)[idx]
        otrace, ptrace = {}, {}
        ocalib.reconstruct(oqm, oblk, unit_id, str(idx), cali, plan=ocalib.DrawPlan(), trace=otrace, **kw)
        Q.block_reconstruction(pqm, pblk, str(idx), cali.to(dev), asym=True, act_quant=False, opt_mode='mse',
                               args=Args(), plan=ReplayPlan(), unit_id=unit_id, trace=ptrace, **kw)
        assert rel_err(ptrace["out"], otrace["out0"]) < 1e-4, (sub, idx)
        assert len(ptrace["d_alpha"]) == len(otrace["grad0"]) >= 2
        for a, b, h0 in zip(ptrace["d_alpha"], otrace["grad0"], otrace["h0_all"]):
            inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
            assert rel_err(a.cpu()[inner], b[inner]) < 2e-3, (sub, idx)
        omods = [m for m in oblk.modules() if isinstance(m, owrap.QuantModule)]
        pmods = [m for m in pblk.modules() if isinstance(m, Q.QuantModule)]
        for a, b, h0 in zip(pmods, omods, otrace["h0_all"]):
            inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
            assert (a.weight_quantizer.alpha.data.cpu() - b.weight_quantizer.alpha.data)[inner].abs().max().item() < 5e-3
        assert pblk.trained and all(m.trained for m in pmods)


def test_reference_style_autograd_loop_still_works(dev):
    """Drop-in check: the reference's own loop shape (AdaRoundQuantizer + torch.optim.Adam + err.backward())."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=8, M=12), 1.2)
    layer = pqm.model.g_a[2]
    (q_in, fp_in), fp_out = Q.save_inp_oup_data(pqm, layer, cali.to(dev), True, False, batch_size=1, input_prob=True)
    layer.weight_quantizer = Q.AdaRoundQuantizer(uaq=layer.weight_quantizer, round_mode='learned_hard_sigmoid',
                                                 weight_tensor=layer.org_weight.data)
    layer.weight_quantizer.soft_targets = True
    layer.set_quant_state(True, False)
    opt = torch.optim.Adam([layer.weight_quantizer.alpha])
    before = Q.lp_loss(layer(q_in), fp_out).item()
    for _ in range(40):
        opt.zero_grad()
        out = layer(q_in)
        _, g = __import__("rdo_ptq_b200").ops.lp_loss_fwd_bwd(out.detach(), fp_out, 2.0, scale=1.0 / (out.numel() // out.shape[1]))
        out.backward(g)
        opt.step()
    assert Q.lp_loss(layer(q_in), fp_out).item() < before


def test_quant_model_pickles(dev, tmp_path):
    """main2.py:285-290 saves the whole QuantModel object."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=8, M=12), 1.2)
    p = tmp_path / "qnn.pth"
    torch.save(pqm, p)
    back = torch.load(p, weights_only=False)
    x = cali[:1].to(dev)
    with torch.no_grad():
        assert torch.equal(back(x)["x_hat"], pqm(x)["x_hat"])


# ---------------------------------------------------------------------------------------------------------------------
# device-resident schedule + CUDA-graph replay of the calibration iteration
# ---------------------------------------------------------------------------------------------------------------------
def test_device_schedule_matches_host_formulas(dev):
    """b200lic_calib_sched_tick reproduces LinearTempDecay (utils.py:37-54), the warm-up gate (layer_opt.py:156-158) and
    Adam's bias corrections for every step of a short schedule."""
    from rdo_ptq_b200 import ops
    from rdo_ptq_b200.quantization import LinearTempDecay
    iters, warmup, b0, b1, lr = 50, 0.2, 20, 2, 1e-3
    decay = LinearTempDecay(iters, rel_start_decay=warmup, start_b=b0, end_b=b1)
    s = ops.new_sched(dev)
    for step in range(1, iters + 1):
        ops.sched_tick(s, iters, warmup, b0, b1, lr)
        got = ops.read_sched(s)
        assert got["step"] == step
        want_b = 0.0 if step < iters * warmup else float(decay(step))
        assert got["reg_b"] == pytest.approx(want_b, rel=1e-6, abs=1e-7)
        be1, be2 = float(torch.tensor(0.9, dtype=torch.float32)), float(torch.tensor(0.999, dtype=torch.float32))
        assert got["lr_over_bc1"] == pytest.approx(float(torch.tensor(lr, dtype=torch.float32)) / (1 - be1 ** step), rel=1e-6)
        assert got["inv_sqrt_bc2"] == pytest.approx(1 / math.sqrt(1 - be2 ** step), rel=1e-6)   # betas are fp32 in the ABI


def test_gather_mix_sched_follows_the_schedule(dev):
    from rdo_ptq_b200 import ops
    g = torch.Generator().manual_seed(3)
    q, fp = torch.randn(6, 5, 4, 4, generator=g).to(dev), torch.randn(6, 5, 4, 4, generator=g).to(dev)
    table = torch.stack([torch.randperm(6, generator=g)[:3] for _ in range(7)]).to(dev)
    s = ops.new_sched(dev)
    units, seed_base = 4, 12345
    for step in range(1, 12):
        ops.sched_tick(s, 100, 0.2, 20, 2)
        for unit in range(units):
            k = (step - 1) * units + unit
            got = ops.gather_mix_sched(q, fp, table, 3, 0.5, seed_base, units, unit, s)
            want = ops.gather_mix(q, fp, table[k % 7].contiguous(), prob=0.5, seed=(seed_base + k) & 0xFFFFFFFFFFFF)
            assert torch.equal(got, want)
            ident = ops.gather_mix_sched(q[:3].contiguous(), fp[:3].contiguous(), None, 3, 1.0, seed_base, units, unit, s)
            assert torch.equal(ident, q[:3])
    # the QDrop draw keeps ~prob of the quantised input and every element comes from one of the two sources
    big_q, big_f = torch.zeros(4, 8, 64, 64, device=dev), torch.ones(4, 8, 64, 64, device=dev)
    mix = ops.gather_mix(big_q, big_f, None, prob=0.5, seed=7)
    assert 0.48 < (mix == 0).float().mean().item() < 0.52
    odd = ops.gather_mix(big_q[:, :, :, :63].contiguous(), big_f[:, :, :, :63].contiguous(), None, prob=0.25, seed=7)
    assert 0.23 < (odd == 0).float().mean().item() < 0.27          # scalar (non-float4) path


@pytest.mark.parametrize("host_caches", [False, True])
def test_session_graph_replay_matches_eager(dev, host_caches):
    """The captured-graph sweep must walk the same trajectory as the eagerly issued one (same device schedule, same
    batch picks and QDrop draws); differences are limited to fp32 atomics ordering in split-K wgrad / loss sums."""
    from rdo_ptq_b200.quantization.session import CalibrationSession

    def run(graph):
        _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
        sess = CalibrationSession(pqm, cali.to(dev), batch_size=2, iters=20, host_caches=host_caches, graph=graph)
        for _ in range(8):
            sess.sweep()
        torch.cuda.synchronize()
        losses = sess.losses()
        alphas = {n: [m.weight_quantizer.alpha.data.clone() for m in t.mods] for n, t in sess.trainers.items()}
        return sess, losses, alphas

    se, le, ae = run(False)
    sg, lg, ag = run(True)
    assert ops_step(se) == ops_step(sg) == 8
    assert len(sg._graphs) == len(sg.units) and sg.replayed_launches > 0
    assert sg.launch_total() > 0
    for n in ae:
        for a, b in zip(ae[n], ag[n]):
            assert (a - b).abs().max().item() < 2e-3, n         # 8 Adam steps of 1e-3; sign flips of ~0 gradients allowed
            assert ((a - b).abs() > 1e-5).float().mean().item() < 0.02, n
        assert lg[n]["rec"] == pytest.approx(le[n]["rec"], rel=1e-3, abs=1e-7), n
        assert lg[n]["round"] == pytest.approx(le[n]["round"], rel=1e-3, abs=1e-7), n


def ops_step(sess):
    from rdo_ptq_b200 import ops
    return ops.read_sched(sess.sched)["step"]


def test_layer_reconstruction_default_path_uses_graph_and_converges(dev):
    """Reference entry point with its default (device-drawn) randomness: graph-replayed loop, loss goes down, the unit is
    hardened afterwards."""
    _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
    layer = pqm.model.g_a[2]

    class Args:
        task_loss = 2.0
    losses = Q.layer_reconstruction(pqm, layer, "2", cali.to(dev), batch_size=2, iters=120, weight=0.01, b_range=(20, 2),
                                    warmup=0.2, input_prob=0.5, asym=True, act_quant=False, opt_mode='mse', args=Args(),
                                    log_every=40)
    assert len(losses) == 3 and losses[-1]["rec"] < losses[0]["rec"]
    assert losses[-1]["round"] > 0.0 and losses[0]["round"] == 0.0 or losses[1]["round"] > 0.0
    assert layer.trained and not layer.weight_quantizer.soft_targets


def test_graphed_evaluation_forward_equals_eager(dev):
    """evaluate.GraphedForward replays the W8A8 forward as one CUDA graph: same x_hat / likelihoods / bits as eager."""
    from rdo_ptq_b200 import evaluate as E
    _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2, bits=8)
    for m in pqm.modules():
        if hasattr(m, "trained"):
            m.trained = True
    pqm.set_quant_state(True, True)
    imgs = [i.to(dev) for i in synth.synthetic_images(3, 128, 192)]
    gf = E.GraphedForward(pqm)
    for k, img in enumerate(imgs):
        xp = E.pad(img, 64)
        with torch.no_grad():
            ref = pqm(xp)
            ref_bits = E.total_bits(ref).item()
        out, bits = gf(xp)                                        # call 0 eager, call 1 captures, call 2 replays
        assert torch.equal(out["x_hat"], ref["x_hat"]), k
        assert torch.equal(out["likelihoods"]["y"], ref["likelihoods"]["y"]), k
        assert abs(bits.item() - ref_bits) < 1e-3 * abs(ref_bits) + 1e-3, k
    assert len(gf.cache) == 1
    res_g = E.evaluate(pqm, imgs, p=64, graph=True)
    res_e = E.evaluate(pqm, imgs, p=64, graph=False)
    assert abs(res_g["psnr"] - res_e["psnr"]) < 1e-6 and abs(res_g["bpp"] - res_e["bpp"]) < 1e-6


def test_streaming_session_matches_cached_session(dev):
    """host_caches="stream" recomputes every unit's (quant_in, fp_in, fp_out) from the batch images with two captured
    forwards; with a pool of exactly one batch the tensors must equal the cached ones and the sweeps must agree."""
    from rdo_ptq_b200.quantization.session import CalibrationSession

    def run(mode):
        _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
        cal = cali[:2]                                              # pool == batch: every pick is a permutation of it
        sess = CalibrationSession(pqm, cal if mode == "stream" else cal.to(dev), batch_size=2, iters=20,
                                  host_caches=mode, input_prob=1.0, n_streams=2)
        return sess

    ss, sc = run("stream"), run(False)
    for n in sc.caches:                                            # streamed tensors == cached tensors (same 2 samples)
        for a, b in zip(ss.caches[n], sc.caches[n]):
            assert torch.allclose(a, b, rtol=0, atol=0) or rel_err(a, b) < 1e-6, n
    for _ in range(6):
        ss.sweep()
        sc.sweep()
    torch.cuda.synchronize()
    ls, lc = ss.losses(), sc.losses()
    assert ss.h2d_bytes == 6 * 4 * 2 * 3 * 64 * 64 and ss._fwd_graph is not None
    for n in lc:
        # the batch is the whole pool in both runs, only its row order differs: the mean loss is order-independent
        assert ls[n]["rec"] == pytest.approx(lc[n]["rec"], rel=2e-3, abs=1e-7), n


def test_w10a10_cheng2020_forward_parity(dev):
    """BASELINE config 4: Cheng2020-attention at W10A10.  The reference asserts n_bits <= 8 and hard-wires 8-bit
    activations; both sides lift that through the same additive switch.  10-bit weight codes bit-exact,
    per-layer outputs within 1e-4, activation codes on the 1023-level grid, bpp / PSNR against the oracle."""
    from rdo_ptq_b200 import evaluate as E
    from rdo_ptq_b200.quantization.quantizer import UniformAffineQuantizer as PUAQ
    wq = dict(n_bits=10, channel_wise=True, scale_method="max")
    aq = dict(n_bits=10, channel_wise=True, scale_method="max", leaf_param=False)
    om, pm, Q = build_pair("cheng2020-attn", dict(N=24), 0.6, dev)
    x = synth.synthetic_image(64, 128)
    with torch.no_grad():
        om(x), pm(x.to(dev))
    oq.UniformAffineQuantizer.act_bits_follow_n_bits = PUAQ.act_bits_follow_n_bits = True
    try:
        oqm, pqm = owrap.QuantModel(om, wq, aq).eval(), Q.QuantModel(pm, wq, aq).eval()
        for q in (oqm, pqm):
            q.set_quant_state(True, False)
        ref, ref_rows = per_layer_io(oqm, x, (owrap.QuantModule,))
        with torch.no_grad():
            pqm(x.to(dev))
        pmods = dict((n, m) for n, m in pqm.named_modules() if isinstance(m, Q.QuantModule))
        omods = dict((n, m) for n, m in oqm.named_modules() if isinstance(m, owrap.QuantModule))
        top = 0.0
        for n, m in pmods.items():
            if m.weight is not None:
                codes = m.weight_quantizer.codes(m.weight).cpu()
                assert torch.equal(codes, omods[n].weight_quantizer.codes(omods[n].weight)), n
                top = max(top, codes.max().item())
        assert 255 < top <= 1023                                   # the 10-bit grid is really in use

        def chk(name, a, b):
            assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))
        layer_local_parity(ref_rows, pmods, dev, chk)
        for q in (oqm, pqm):
            for m in q.modules():
                if hasattr(m, "trained"):
                    m.trained = True
            q.set_quant_state(True, True)
        ref10, rows10 = per_layer_io(oqm, x, (owrap.QuantModule,))
        with torch.no_grad():
            out10 = pqm(x.to(dev))

        def chk10(name, a, b):
            step = (b.amax() - b.amin()).item() / 1023 + 1e-12
            assert ((a - b).abs() <= 1.01 * step).all(), name
        layer_local_parity(rows10, pmods, dev, chk10)
        bpp_ref, bpp = oeval.compute_bpp(ref10), E.compute_bpp(out10)
        p_ref = oeval.compute_psnr(x, ref10["x_hat"].clamp(0, 1))
        p = E.compute_psnr(out10["x_hat"], x.to(dev), clamp=True)
        assert abs(bpp - bpp_ref) < 0.01 * bpp_ref + 1e-3 and abs(p - p_ref) < 0.1
    finally:
        oq.UniformAffineQuantizer.act_bits_follow_n_bits = PUAQ.act_bits_follow_n_bits = False


def test_main2_whole_model_calibration_matches_oracle(dev):
    """The calibration entry point (main2.py:145-290): QuantModel -> 8-bit head/stem -> range init -> depth-first
    recon_model walk (every unit sees the hardened units before it) -> W4 test -> W4A8 test, against the oracle's
    restatement of the same walk on the same draws.  4-bit weights so that rounding decisions matter."""
    from rdo_ptq_b200 import main2
    argv = ["--arch", "Minnen2018", "--n_bits_w", "4", "--channel_wise", "--batch_size", "2", "--num_samples", "4",
            "--iters_w", "12", "--test_before_calibration"]
    args = main2.parse_args(argv)
    om, pm, Q = build_pair("mbt2018-mean", dict(N=16, M=24), 1.2, dev)
    cali = synth.calibration_patches(4, 64)
    imgs = synth.synthetic_images(2, 100, 150)
    # oracle side, statement by statement
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
    oqm.model.g_s[-1].set_quant_state(True, False)
    otr = ocalib.recon_model(oqm, cali, batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2,
                             input_prob=0.5, plan=ocalib.DrawPlan())
    oqm.set_quant_state(True, False)
    ps_w, bs_w = oeval.evaluate(oqm, imgs)
    oqm.set_quant_state(True, True)
    oqm.model.g_s[-1].set_quant_state(True, False)
    ps_wa, bs_wa = oeval.evaluate(oqm, imgs)
    # product side: the entry point
    pqm, rep = main2.optimize_model(args, model=pm, cali_data=cali, test_images=imgs, device=dev, plan=ReplayPlan())
    assert list(rep["losses"]) == list(otr) and len(otr) == 20
    omods = [m for m in oqm.modules() if isinstance(m, owrap.QuantModule) and m.org_weight is not None]
    pmods = [m for m in pqm.modules() if isinstance(m, Q.QuantModule) and m.org_weight is not None]
    assert len(omods) == len(pmods) == 20 and all(m.trained for m in pmods)
    assert pmods[0].weight_quantizer.n_bits == 8 and pmods[1].weight_quantizer.n_bits == 4
    diff = tot = 0
    for a, b in zip(pmods, omods):
        assert torch.equal(a.weight_quantizer.delta.cpu().reshape(-1), b.weight_quantizer.delta.reshape(-1))
        ca, cb = a.weight_quantizer.codes(a.weight).cpu(), b.weight_quantizer.codes(b.weight)
        diff += (ca != cb).sum().item()
        tot += ca.numel()
    print(f"main2 walk: {diff}/{tot} hardened 4-bit codes differ; W4 bpp {rep['w_opt']['bpp']:.5f} vs "
          f"{sum(bs_w) / 2:.5f}, psnr {rep['w_opt']['psnr']:.4f} vs {sum(ps_w) / 2:.4f}; W4A8 bpp "
          f"{rep['wa_opt']['bpp']:.5f} vs {sum(bs_wa) / 2:.5f}, psnr {rep['wa_opt']['psnr']:.4f} vs {sum(ps_wa) / 2:.4f}")
    assert diff / tot < 2e-3                   # alpha within float noise of 0 after 12 steps: isolated decisions only
    # hardened units, layer by layer on the oracle's own inputs: the per-layer bar (1e-4)
    oqm.set_quant_state(True, False)
    pqm.set_quant_state(True, False)
    _, rows = per_layer_io(oqm, oeval.pad(imgs[0], 256), (owrap.QuantModule,))
    pm_by_name = dict((n, m) for n, m in pqm.named_modules() if isinstance(m, Q.QuantModule))

    def chk(name, a, b):
        assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))
    layer_local_parity(rows, pm_by_name, dev, chk)
    # End to end this random-init 4-bit codec is chaotic (PSNR ~ 6 dB): one latent symbol that rounds the other way
    # changes every mean / scale after it, so the metrics of two correct implementations agree to a few percent only
    # (the W8 end-to-end bars of 1e-3 bpp / 0.01 dB are checked in test_quantized_forward_per_layer_parity).
    # (observed on B200: W4 bpp 0.1501 vs 0.1587, PSNR 5.786 vs 5.760 dB with every hardened code identical)
    assert abs(rep["w_opt"]["bpp"] - sum(bs_w) / 2) < 0.12 * sum(bs_w) / 2
    assert abs(rep["w_opt"]["psnr"] - sum(ps_w) / 2) < 0.3
    assert abs(rep["wa_opt"]["bpp"] - sum(bs_wa) / 2) < 0.12 * sum(bs_wa) / 2
    assert abs(rep["wa_opt"]["psnr"] - sum(ps_wa) / 2) < 0.3
    assert "fp32" in rep and "w_nearest" in rep


@pytest.mark.parametrize("init", ["mse", "gaussian"])
def test_main2_default_path_with_search_init(dev, init):
    """Entry point with its own (device-drawn, graph-replayed) randomness and a search-based --init."""
    from rdo_ptq_b200 import main2
    args = main2.parse_args(["--arch", "mbt2018-mean", "--N", "8", "--M", "12", "--n_bits_w", "4", "--channel_wise",
                             "--batch_size", "2", "--num_samples", "4", "--iters_w", "24", "--patch", "64", "--init",
                             init, "--test_hw", "64x96", "--n_test", "1"])
    qnn, rep = main2.optimize_model(args, device=dev)
    assert len(rep["losses"]) == 20 and all(len(v) >= 1 for v in rep["losses"].values())
    assert math.isfinite(rep["wa_opt"]["bpp"]) and math.isfinite(rep["wa_opt"]["psnr"])
    assert all(m.weight_quantizer.scale_method == init if hasattr(m.weight_quantizer, "scale_method") else True
               for m in qnn.modules() if hasattr(m, "weight_quantizer"))


def test_lu_quantize_entry_point(dev, tmp_path):
    """light-uniform-PTQ/quantize.py flow (config 1): uint8 weights materialised by one forward, INT8 state dict saved."""
    from rdo_ptq_b200 import quantize as lu_entry, quant_int as LU
    p = tmp_path / "INT8.pth"
    args = lu_entry.parse_args(["--N", "8", "--M", "12", "--hw", "64x96", "--n_test", "2", "--save", str(p)])
    qnn, rep = lu_entry.quantize_int8(args, device=dev)
    mods = [m for m in qnn.modules() if isinstance(m, LU.QuantModule)]
    assert len(mods) == 14 and all(m.weight.dtype == torch.uint8 for m in mods)
    assert abs(rep["int8"]["psnr"] - rep["fp32"]["psnr"]) < 1.0 and rep["int8"]["bpp"] > 0
    sd = torch.load(p, weights_only=False)
    assert any(v.dtype == torch.uint8 for v in sd.values())


def test_lagged_loss_readback_returns_the_previous_window(dev):
    """CalibrationSession.losses(lag=True): non-blocking copy into pinned memory, values of the previous call's window."""
    from rdo_ptq_b200.quantization.session import CalibrationSession

    def make():
        _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
        return CalibrationSession(pqm, cali.to(dev), batch_size=2, iters=20)
    a, b = make(), make()
    ref, got = [], []
    for _ in range(5):
        a.sweep()
        ref.append(a.losses())
        b.sweep()
        got.append(b.losses(lag=True))
    assert got[0] == {}
    for k in range(1, 5):
        for n in ref[k - 1]:
            # (block partial sums meet in fp32 atomics: the order, hence the last bits, can differ between two runs)
            assert got[k][n]["rec"] == pytest.approx(ref[k - 1][n]["rec"], rel=1e-4, abs=1e-9), (k, n)
            assert got[k][n]["round"] == pytest.approx(ref[k - 1][n]["round"], rel=1e-4, abs=1e-9), (k, n)
