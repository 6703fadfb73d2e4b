"""GPU parity of the rate-side backward kernels (b200lic_gaussian_lik_bwd, b200lic_factorized_lik_bwd), the learned
step-size gradient (b200lic_lsq_delta_grad), the differentiable R + lambda*D loss and the R + lambda*D task criterion
of the calibration loop, each against torch autograd of the CPU oracle on the same seeded inputs.

Tolerance: 1e-4 relative (norm) on floating-point gradients, stated per assertion."""
import math

import pytest
import torch

from oracle import codec as ocodec, quantizers as oq, calib as ocalib, evalpath as oeval
from rdo_ptq_b200 import synth

from test_gpu_model import _calib_pair, ReplayPlan, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(dev):
    from rdo_ptq_b200 import ops as _ops
    return _ops


# --------------------------------------------------------------------------------------------- K9 backward
@pytest.mark.parametrize("ste", [False, True])
@pytest.mark.parametrize("with_means", [True, False])
@pytest.mark.parametrize("shape", [(2, 6, 5, 7), (3, 8, 8, 12), (1, 320, 32, 48)])
def test_gaussian_likelihood_backward(ops, dev, ste, with_means, shape):
    gen = torch.Generator().manual_seed(7)
    y = (torch.randn(shape, generator=gen) * 3).requires_grad_(True)
    gp = torch.randn(shape[0], 2 * shape[1], *shape[2:], generator=gen)
    gp[:, :shape[1]] = gp[:, :shape[1]].abs() * 1.5 + 0.02           # scales: some below the 0.11 bound (LowerBound gate)
    gp.requires_grad_(True)
    G = torch.randn(shape, generator=gen)                             # upstream dL/dlik, random sign (likelihood gate)
    H = torch.randn(shape, generator=gen)                             # upstream dL/dy_hat
    c = 0.37                                                          # dL/dbits
    gc = ocodec.GaussianConditional(None).eval()
    gc.ste_round = ste
    sc, mu = gp.chunk(2, 1)
    yh, lik = gc(y, sc, means=mu if with_means else None)
    loss = (lik * G).sum() + c * (-torch.log2(lik)).sum() + (yh * H).sum()
    loss.backward()

    yd = y.detach().to(dev).requires_grad_(True)
    gpd = gp.detach().to(dev).requires_grad_(True)
    scd, mud = gpd.chunk(2, 1)                                        # strided parameter views, as in the codec
    yh2, lik2, bits2 = ops.gaussian_lik_fn(yd, scd, mud if with_means else None, ste=ste)
    assert torch.equal(yh2.detach().cpu(), yh.detach())
    loss2 = (lik2 * G.to(dev)).sum() + c * bits2.sum() + (yh2 * H.to(dev)).sum()
    loss2.backward()
    assert abs(loss2.item() - loss.item()) < 1e-4 * abs(loss.item()) + 1e-2
    assert rel_err(gpd.grad, gp.grad) < 1e-4, rel_err(gpd.grad, gp.grad)
    if ste:
        assert rel_err(yd.grad, y.grad) < 1e-4
    else:                                                             # torch.round: no gradient reaches y
        assert y.grad is None or float(y.grad.abs().max()) == 0.0
        assert yd.grad is None or float(yd.grad.abs().max()) == 0.0


def test_gaussian_backward_bound_gates(ops, dev):
    """LowerBound semantics on both bounds: a likelihood clamped at 1e-9 passes only gradients that push it up, a scale
    clamped at 0.11 only gradients that push it up (compressai LowerBound; oracle codec._LowerBoundFn)."""
    y = torch.tensor([[[[0.0, 40.0, 40.0, 0.2, 0.2]]]])
    sc = torch.tensor([[[[0.05, 0.5, 0.5, 0.05, 0.05]]]], requires_grad=True)       # 0.05 < bound
    G = torch.tensor([[[[1.0, 1.0, -1.0, 1.0, -1.0]]]])
    gc = ocodec.GaussianConditional(None).eval()
    _, lik = gc(y, sc)
    (lik * G).sum().backward()
    scd = sc.detach().to(dev).requires_grad_(True)
    _, lik2, _ = ops.gaussian_lik_fn(y.to(dev), scd, None)
    (lik2 * G.to(dev)).sum().backward()
    assert torch.allclose(scd.grad.cpu(), sc.grad, rtol=1e-4, atol=1e-12), (scd.grad, sc.grad)
    assert float(sc.grad[0, 0, 0, 1]) == 0.0 and float(scd.grad[0, 0, 0, 1]) == 0.0   # clamped lik, g > 0: blocked


# --------------------------------------------------------------------------------------------- K10 backward
@pytest.mark.parametrize("shape", [(2, 5, 6, 7), (2, 192, 8, 12)])
def test_factorized_likelihood_backward(ops, dev, shape, golden_c):
    from rdo_ptq_b200.codec import EntropyBottleneck
    torch.manual_seed(3)
    C = shape[1]
    oe = ocodec.EntropyBottleneck(C).eval()
    if C == 5:
        oe.load_state_dict(golden_c["eb"]["state"])                    # trained-looking parameters (non-zero factors)
    else:
        with torch.no_grad():
            for i in range(4):
                getattr(oe, f"_factor{i}").normal_(0, 0.5)
    pe = EntropyBottleneck(C).eval()
    pe.load_state_dict(oe.state_dict())
    pe.to(dev)
    gen = torch.Generator().manual_seed(11)
    z = (torch.randn(shape, generator=gen) * 4).requires_grad_(True)
    G, H, c = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen), 0.21
    for ste in (True, False):
        oe.ste_round = pe.ste_round = ste
        z.grad = None
        zh, lik = oe(z)
        loss = (lik * G).sum() + c * (-torch.log2(lik)).sum() + (zh * H).sum()
        loss.backward()
        zd = z.detach().to(dev).requires_grad_(True)
        zh2, lik2 = pe(zd)
        loss2 = (lik2 * G.to(dev)).sum() + c * pe.last_bits.sum() + (zh2 * H.to(dev)).sum()
        loss2.backward()
        assert torch.equal(zh2.detach().cpu(), zh.detach())
        if ste:
            assert rel_err(zd.grad, z.grad) < 1e-4, rel_err(zd.grad, z.grad)
        else:
            assert float(zd.grad.abs().max()) == 0.0 and float(z.grad.abs().max()) == 0.0


# --------------------------------------------------------------------------------------------- learned step size
@pytest.mark.parametrize("wshape,tconv", [((24, 16, 5, 5), False), ((16, 24, 5, 5), True), ((12, 12), False)])
def test_lsq_delta_gradient_matches_autograd(ops, dev, wshape, tconv):
    """d loss / d delta by autograd of the oracle's fake-quant expressions (quantizer.py:175-177 and :437-449) with
    delta as the leaf, against b200lic_lsq_delta_grad; then three fused Adam steps on delta against torch.optim.Adam."""
    gen = torch.Generator().manual_seed(21)
    w = torch.randn(wshape, generator=gen) * 0.1
    g = torch.randn(wshape, generator=gen)
    uq = oq.UniformAffineQuantizer(n_bits=4, channel_wise=True, scale_method="max", tconv=tconv)
    uq(w)                                                              # initialises delta / zero_point
    axis = oq.channel_axis(w.shape, tconv)
    delta0 = (uq.delta * 0.9).clone()                                  # off the min/max point so clamping is live
    zp = uq.zero_point
    # nearest rounding
    d = delta0.clone().requires_grad_(True)
    x_int = oq.round_ste(w / d) + zp
    out = (torch.clamp(x_int, 0, uq.n_levels - 1) - zp) * d
    (out * g).sum().backward()
    ref = d.grad.reshape(-1)
    got = ops.lsq_delta_grad(w.to(dev), delta0.to(dev), zp.to(dev), g.to(dev), axis, uq.n_levels)
    assert rel_err(got, ref) < 1e-4, rel_err(got, ref)
    # AdaRound soft forward (floor has no gradient)
    uq.delta = delta0.clone()
    ar = oq.AdaRoundQuantizer(uq, w)
    ar.soft_targets = True
    ar.delta = delta0.clone().requires_grad_(True)
    (ar(w) * g).sum().backward()
    ref = ar.delta.grad.reshape(-1)
    got = ops.lsq_delta_grad(w.to(dev), delta0.to(dev), zp.to(dev), g.to(dev), axis, uq.n_levels,
                             alpha=ar.alpha.detach().to(dev), soft=True)
    assert rel_err(got, ref) < 1e-4, rel_err(got, ref)
    # fused Adam on delta
    d = delta0.clone().requires_grad_(True)
    opt = torch.optim.Adam([d], lr=1e-4)
    dd = delta0.to(dev).clone()
    m, v = torch.zeros(dd.numel(), device=dev), torch.zeros(dd.numel(), device=dev)
    for step in range(1, 4):
        opt.zero_grad()
        out = (torch.clamp(oq.round_ste(w / d) + zp, 0, uq.n_levels - 1) - zp) * d
        (out * g).sum().backward()
        opt.step()
        ops.lsq_delta_grad(w.to(dev), dd, zp.to(dev), g.to(dev), axis, uq.n_levels, adam=(m, v, step, 1e-4))
    assert rel_err(dd, d.detach()) < 1e-5


# --------------------------------------------------------------------------------------------- R + lambda*D
def _rd_models(dev, arch="mbt2018-mean", kw=dict(N=16, M=24)):
    from test_gpu_model import build_pair
    om, pm, _ = build_pair(arch, kw, 1.2, dev)
    for m in (om, pm):
        m.entropy_bottleneck.ste_round = m.gaussian_conditional.ste_round = True
    return om, pm


@pytest.mark.parametrize("arch", ["mbt2018-mean", "bmshj2018-hyperprior"])
@pytest.mark.parametrize("coder", ["g_a", "h_a", "h_s", "g_s"])
def test_rd_loss_gradient_through_the_codec_tail(ops, dev, arch, coder):
    """loss = lambda*255^2*MSE + bpp on forward_from(coder, value): value and d loss / d value against oracle autograd
    (conv/deconv dgrad, GDN backward, both likelihood backward kernels, straight-through latent rounding)."""
    om, pm = _rd_models(dev, arch)
    x = synth.calibration_patches(2, 64)
    with torch.no_grad():
        y, z = om.latents(x)
        value = {"g_a": y, "h_a": z, "h_s": om.h_s(om.entropy_bottleneck(z)[0]), "g_s": om(x)["x_hat"]}[coder]
    value = (value + 0.05 * torch.randn(value.shape, generator=torch.Generator().manual_seed(2))).detach()
    lm = 0.013
    v = value.clone().requires_grad_(True)
    o = om.forward_from(coder, v, {"y": y, "z": z})
    ref = oeval.rate_distortion_loss(o, x, lm)["loss"]
    ref.backward()
    vd = value.to(dev).requires_grad_(True)
    o2 = pm.forward_from(coder, vd, {"y": y.to(dev), "z": z.to(dev)})
    loss = ops.rd_loss(o2["x_hat"], x.to(dev), o2["bits"], lm)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    assert rel_err(vd.grad, v.grad) < 2e-4, rel_err(vd.grad, v.grad)


@pytest.mark.parametrize("layer_path", ["g_a.2", "h_a.0", "h_s.2", "g_s.4"])
def test_layer_reconstruction_with_rd_task(dev, layer_path):
    """AdaRound reconstruction with task = R + lambda*D: first-iteration dL/dalpha, task value and the alpha trajectory
    against the oracle loop with the same draws."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
    sub, idx = layer_path.split(".")
    olayer, player = getattr(oqm.model, sub)[int(idx)], getattr(pqm.model, sub)[int(idx)]
    kw = dict(batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)
    lm = 0.01
    otrace, ptrace = {}, {}
    oqm.set_quant_state(False, False)
    ord_ = ocalib.RDTask(oqm, layer_path, cali, lm)
    olosses = ocalib.reconstruct(oqm, olayer, 3, idx, cali, plan=ocalib.DrawPlan(), trace=otrace, rd_task=ord_, **kw)
    ord_.close()
    plosses = Q.layer_reconstruction(pqm, player, idx, cali.to(dev), asym=True, act_quant=False, opt_mode='mse',
                                     plan=ReplayPlan(), unit_id=3, trace=ptrace, task="rd", lmbda=lm,
                                     unit_path=layer_path, log_every=1, **kw)
    assert rel_err(ptrace["out"], otrace["out0"]) < 1e-4
    h0 = otrace["h0"]
    inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
    assert rel_err(ptrace["d_alpha"][0].cpu()[inner], otrace["grad0"][0][inner]) < 1e-3
    assert abs(plosses[0]["total"] - olosses[0]) < 1e-3 * abs(olosses[0]), (plosses[0], olosses[0])
    a_ref, a_gpu = olayer.weight_quantizer.alpha.data, player.weight_quantizer.alpha.data.cpu()
    assert (a_ref - a_gpu)[inner].abs().max().item() < 5e-3
    assert not pqm.model.entropy_bottleneck.ste_round and not pqm.model.gaussian_conditional.ste_round


@pytest.mark.parametrize("layer_path", ["g_a.2", "g_s.2"])
def test_layer_reconstruction_with_learned_step_size(dev, layer_path):
    """AdaRound + learned per-channel delta (the option the reference keeps commented out, layer_opt.py:259-265):
    first-iteration d loss / d delta and the delta / alpha trajectories against the oracle loop."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
    sub, idx = layer_path.split(".")
    olayer, player = getattr(oqm.model, sub)[int(idx)], getattr(pqm.model, sub)[int(idx)]
    kw = dict(batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)
    otrace, ptrace = {}, {}
    ocalib.reconstruct(oqm, olayer, 3, idx, cali, plan=ocalib.DrawPlan(), trace=otrace, learn_delta=True, **kw)

    class Args:
        task_loss = 2.0
    delta_before = player.weight_quantizer.delta.clone()
    Q.layer_reconstruction(pqm, player, idx, cali.to(dev), asym=True, act_quant=False, opt_mode='mse', args=Args(),
                           plan=ReplayPlan(), unit_id=3, trace=ptrace, learn_delta=True, **kw)
    assert rel_err(ptrace["d_delta"][0], otrace["d_delta0"][0]) < 1e-3
    d_ref = olayer.weight_quantizer.delta.detach().reshape(-1)
    d_gpu = player.weight_quantizer.delta.reshape(-1).cpu()
    moved = (d_gpu - delta_before.reshape(-1).cpu()).abs().max().item()
    assert moved > 0.0                                                     # delta was actually trained
    # Adam moves delta by ~lr_delta = 1e-4 per step whatever the gradient's size, and floor(w/delta) makes the loss
    # piecewise in delta: a channel whose tiny gradient flips sign once ends two steps apart.  Bar: every channel within
    # two steps, the typical channel within a hundredth of a step.
    diff = (d_gpu - d_ref).abs()
    assert diff.max().item() < 2.5e-4 and diff.median().item() < 1e-6, (diff.max().item(), diff.median().item())
    h0 = otrace["h0"]
    inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
    a_ref, a_gpu = olayer.weight_quantizer.alpha.data, player.weight_quantizer.alpha.data.cpu()
    assert (a_ref - a_gpu)[inner].abs().max().item() < 5e-3


def test_session_with_learned_step_size_graph_matches_eager(dev):
    from rdo_ptq_b200.quantization.session import CalibrationSession

    def run(graph):
        _, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
        sess = CalibrationSession(pqm, cali.to(dev), batch_size=2, iters=20, graph=graph, learn_delta=True)
        for _ in range(6):
            sess.sweep()
        torch.cuda.synchronize()
        return {n: [m.weight_quantizer.delta.clone() for m in t.mods] for n, t in sess.trainers.items()}

    de, dg = run(False), run(True)
    for n in de:
        for a, b in zip(de[n], dg[n]):
            assert rel_err(b, a) < 1e-4, n


@pytest.mark.parametrize("layer_path", ["g_a.2", "g_s.2", "h_s.0"])
def test_layer_reconstruction_with_coder_task(dev, layer_path):
    """task = lp_loss(tail(out_quant), tail(out_fp)) over the later modules of the unit's own sub-network (+ round_ste
    for g_a): the reference's fp_out rule (layer_opt.py:45-75) applied by position."""
    oqm, pqm, Q, cali = _calib_pair(dev, "mbt2018-mean", dict(N=16, M=24), 1.2)
    sub, idx = layer_path.split(".")
    olayer, player = getattr(oqm.model, sub)[int(idx)], getattr(pqm.model, sub)[int(idx)]
    kw = dict(batch_size=2, iters=12, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)
    otrace, ptrace = {}, {}
    # the oracle's target needs the FP unit outputs: the same cache `reconstruct` builds
    (_, _), fp_out = ocalib.save_inp_oup_data(oqm, olayer, cali, False, False)
    oqm.set_quant_state(False, False)
    otask = ocalib.CoderTask(oqm, layer_path, fp_out, 2.0)
    olosses = ocalib.reconstruct(oqm, olayer, 3, idx, cali, plan=ocalib.DrawPlan(), trace=otrace, rd_task=otask, **kw)

    class Args:
        task_loss = 2.0
    plosses = Q.layer_reconstruction(pqm, player, idx, cali.to(dev), asym=True, act_quant=False, opt_mode='mse',
                                     args=Args(), plan=ReplayPlan(), unit_id=3, trace=ptrace, task="coder",
                                     unit_path=layer_path, log_every=1, **kw)
    h0 = otrace["h0"]
    inner = (h0 > 1e-4) & (h0 < 1 - 1e-4)
    assert rel_err(ptrace["d_alpha"][0].cpu()[inner], otrace["grad0"][0][inner]) < 1e-3
    assert abs(plosses[0]["total"] - olosses[0]) < 1e-3 * abs(olosses[0]) + 1e-7, (plosses[0], olosses[0])
    a_ref, a_gpu = olayer.weight_quantizer.alpha.data, player.weight_quantizer.alpha.data.cpu()
    assert (a_ref - a_gpu)[inner].abs().max().item() < 5e-3


def test_main2_entry_point_with_rd_task_loss(dev):
    """`main2.py --task_loss rd` end to end: recon_model hands every unit its path inside the codec and
    --lmbda, so each of the 20 layer problems runs with task = R + lambda*D; a unit nested inside a composite child
    (no positional tail) raises instead of silently falling back to the lp task."""
    from rdo_ptq_b200 import main2
    from rdo_ptq_b200.quantization import layer_opt
    args = main2.parse_args(["--arch", "mbt2018-mean", "--N", "8", "--M", "12", "--n_bits_w", "4", "--channel_wise",
                             "--batch_size", "2", "--num_samples", "4", "--iters_w", "6", "--patch", "64",
                             "--task_loss", "rd", "--lmbda", "0.01", "--test_hw", "64x96", "--n_test", "1"])
    assert args.task_loss == "rd"
    qnn, rep = main2.optimize_model(args, device=dev)
    assert len(rep["losses"]) == 20
    for name, recs in rep["losses"].items():
        assert recs and recs[-1]["task"] > 0 and math.isfinite(recs[-1]["total"]), name
        # the task term is a rate-distortion value (bits per pixel + lambda * 255^2 * MSE), not a copy of rec
        assert recs[-1]["task"] != recs[-1]["rec"], name
    assert math.isfinite(rep["wa_opt"]["bpp"])
    with pytest.raises(NotImplementedError):
        layer_opt._rd_task(qnn, "g_a.3.conv_a.0", None, args, None, 0.01)
