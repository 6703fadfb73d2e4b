"""GPU: the prepared-operand entry points (include/b200lic.h, "Prepared operands") against the un-fused kernels they
replace.  Every comparison is BIT-EXACT (`torch.equal`): the fused kernels restate the same arithmetic in the same
order, they only skip the round trips through HBM -- so all parity results of the un-fused path carry over."""
import ctypes as C

import pytest
import torch

from rdo_ptq_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(dev):
    from rdo_ptq_b200 import ops as _ops
    return _ops


def _layer(ops, dev, transposed, N=2, Cin=64, Cout=96, H=20, W=24, k=5, st=2, act=0, seed=0):
    g = torch.Generator().manual_seed(1005 + seed)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
    w = (torch.randn(wshape, generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, st, k // 2, transposed, st - 1 if transposed else 0, act=act, slope=0.01)
    return x, w, b, d


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("mode", ["nearest", "soft", "hard", "integer"])
@pytest.mark.parametrize("Cout", [96, 40])              # 40: output channels padded to 48 in the packed operand
def test_quant_pack_equals_quantiser_then_pack(ops, dev, transposed, mode, Cout):
    x, w, b, d = _layer(ops, dev, transposed, Cout=Cout)
    axis = 1 if transposed else 0
    delta, zp = ops.wq_init_minmax(w, axis, 8)
    alpha = ops.adaround_init_alpha(w, delta, axis) + torch.randn_like(w)
    if mode == "nearest":
        ref_w = ops.wq_fake_quant(w, delta, zp, axis, 256)
        packed = ops.quant_pack_weights(w, None, delta, zp, axis, 256, False, d, transposed)
    elif mode == "integer":
        ref_w = ops.wq_int_weights(w, delta, zp, axis, 256, alpha)
        packed = ops.quant_pack_weights(w, alpha, delta, zp, axis, 256, False, d, transposed, integer_mode=True)
    else:
        ref_w = ops.adaround_fwd(w, alpha, delta, zp, axis, 256, mode == "soft")
        w_q = torch.empty_like(w)
        packed = ops.quant_pack_weights(w, alpha, delta, zp, axis, 256, mode == "soft", d, transposed, w_q=w_q)
        assert torch.equal(w_q, ref_w)
    ref = ops.pack_weights(ref_w, d, transposed)
    n = packed.numel() // 2 if mode == "integer" else packed.numel()      # integer mode: hi slab only
    assert torch.equal(packed[:n], ref[:n])
    # and the forward through the prepared operand equals the plain forward on the quantised weight
    if mode == "integer":
        scale = delta.reshape(-1).contiguous()
        y = ops.conv_fwd_packed(x, packed, d, transposed, bias=b, w_scale=scale)
        y_ref = ops.conv_wq(x, ref_w, scale, b, stride=d.stride, padding=d.pad, output_padding=d.stride - 1 if transposed
                            else 0, transposed=transposed)
    else:
        y = ops.conv_fwd_packed(x, packed, d, transposed, bias=b)
        y_ref = ops.deconv2d_raw(x, ref_w, b, d) if transposed else ops.conv2d_raw(x, ref_w, b, d)
    assert torch.equal(y, y_ref)


def test_gdn_forward_with_prepared_gamma(ops, dev):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 64, 16, 24, generator=g).to(dev)
    gam = (torch.rand(64, 64, generator=g) * 0.01 + 0.1 * torch.eye(64)).to(dev)
    bet = (1 + torch.rand(64, generator=g)).to(dev)
    for inverse in (False, True):
        d = ops.gdn_desc(x.shape, inverse)
        y_ref, n_ref = ops.conv2d_raw(x, gam.view(64, 64, 1, 1), bet, d, gdn_x=x, want_norm=True)
        packed = ops.pack_weights(gam.view(64, 64, 1, 1), d, False)
        y, n = ops.conv_fwd_packed(x, packed, d, False, bias=bet, gdn_x=x, want_norm=True)
        assert torch.equal(y, y_ref) and torch.equal(n, n_ref)


@pytest.mark.parametrize("shape", [(2, 64, 16, 24), (1, 192, 64, 96), (3, 192, 30, 22), (1, 128, 200, 130), (2, 80, 12, 20),
                                   (1, 256, 40, 36), (148, 192, 16, 16)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("deferred", [False, True])
def test_gdn_one_kernel_forward(ops, dev, shape, inverse, deferred):
    """b200lic_gdn_fwd_fused (operand squared / split / quantised on chip, MN-major tcgen05 operand) against the staged
    conv-engine form: fp64 GDN as the yardstick for both, and the two against each other at fp32 rounding level.
    Ragged pixel tails (HW % 128 != 0), several images, channel counts that pad to 32, more tiles than SMs."""
    N, Cc, H, W = shape
    g = torch.Generator().manual_seed(11 + Cc + H)
    x = (torch.randn(N, Cc, H, W, generator=g) * 2).to(dev)
    gam = (torch.rand(Cc, Cc, generator=g) * 0.02 + 0.1 * torch.eye(Cc)).to(dev)
    bet = (1 + torch.rand(Cc, generator=g)).to(dev)
    d = ops.gdn_desc(x.shape, inverse)
    packed = ops.pack_weights(gam.view(Cc, Cc, 1, 1), d, False)
    assert ops.gdn_fused_ok(Cc, H * W)
    if deferred:
        keys = ops.act_quant_stats(x)
        xq = ops.act_quant_apply(x, keys, 8)
        y = ops.gdn_fwd_fused(x, packed, bet, inverse, pending=(keys, 8))
    else:
        xq = x
        y = ops.gdn_fwd_fused(x, packed, bet, inverse)
    y_ref = ops.conv_fwd_packed(xq, packed, d, False, bias=bet, gdn_x=xq)
    x64 = xq.double()
    nrm = torch.einsum("ok,nkhw->nohw", gam.double(), x64 * x64) + bet.double().view(1, -1, 1, 1)
    y64 = x64 * (nrm.sqrt() if inverse else nrm.rsqrt())
    scale = y64.abs().max().item()
    e_new = (y.double() - y64).abs().max().item() / scale
    e_old = (y_ref.double() - y64).abs().max().item() / scale
    assert e_new < 5e-6 and e_new < 2 * e_old + 2e-7, (e_new, e_old)
    assert (y - y_ref).abs().max().item() / scale < 5e-6      # one flipped 8-bit code would be ~4e-3


def test_gdn_one_kernel_division_is_correctly_rounded(ops, dev):
    """The on-chip quantiser divides with an FMA sequence on a precomputed reciprocal; it must round like the IEEE
    division of b200lic_actq_apply for the codes to stay bit-identical: 2^28 pseudo-random (value, range) pairs and the
    exhaustive code / levels tables of 2..16 bit."""
    from rdo_ptq_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.call("selftest_fast_div", 1 << 28, 1005, bad.data_ptr())
    assert int(bad.item()) == 0


def test_gdn_one_kernel_rejects_what_it_cannot_tile(ops, dev):
    assert not ops.gdn_fused_ok(192, 30 * 31 + 1)      # HW % 4 != 0: TMA row pitch
    assert not ops.gdn_fused_ok(320, 64 * 64)          # accumulator wider than 256 TMEM columns


@pytest.mark.parametrize("transposed,Cin", [(False, 64), (True, 64), (False, 96)])
@pytest.mark.parametrize("prob", [0.5, 1.0])
def test_stage_mix_equals_gather_mix_then_split(ops, dev, transposed, Cin, prob):
    """pick + QDrop straight into the staged operand == gather_mix_sched -> fp32 batch -> NHWC split (forward and
    weight gradient consume the staged operand; 96 channels: the forward pads to 96, the wgrad engine to 128)."""
    x, w, b, d = _layer(ops, dev, transposed, N=4, Cin=Cin)
    g = torch.Generator().manual_seed(3)
    pool_q = torch.randn(6, *x.shape[1:], generator=g).to(dev)
    pool_f = torch.randn(6, *x.shape[1:], generator=g).to(dev)
    table = torch.stack([torch.randperm(6, generator=g)[:4] for _ in range(5)]).to(dev)
    sched = ops.new_sched(dev)
    for _ in range(3):
        ops.sched_tick(sched, 100, 0.2, 20, 2)
    cur = ops.gather_mix_sched(pool_q, pool_f, table, 4, prob, 12345, 7, 2, sched)
    y_ref, fwd_ws = (ops.deconv2d_raw if transposed else ops.conv2d_raw)(cur, w, b, d, want_ws=True)
    ws = ops._workspace(d, ops.fwd_op(transposed), dev)
    slot = ops.conv_x_slot(d, transposed, ws)
    assert slot is not None and slot[2] == (Cin + 31) // 32 * 32
    out = torch.empty_like(cur)
    ops.stage_mix_sched(pool_q, pool_f, table, 4, prob, 12345, 7, 2, sched, slot, out=out)
    assert torch.equal(out, cur)
    packed = ops.pack_weights(w, d, transposed)
    y = ops.conv_fwd_packed(None, packed, d, transposed, bias=b, ws=ws)
    assert torch.equal(y, y_ref)
    # weight gradient from the staged x (channel pitch = the forward's padding) == the plain weight gradient
    dy = torch.randn(y.shape, generator=g).to(dev)
    dw_ref = torch.empty_like(w)
    ops._wgrad(d, transposed, cur, dy, dw_ref, (None, 0))
    ws_w = ops._workspace(d, ops.wgrad_op(transposed), dev)
    dw = torch.empty_like(w)
    ops.conv_wgrad_prepared(d, transposed, slot, dy, dw, ws_w)
    assert torch.equal(dw, dw_ref)


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("act", [0, 2])
def test_loss_stage_and_fused_tail_equal_the_unfused_chain(ops, dev, transposed, act):
    """loss + gradient -> staged dY -> wgrad -> (fused) STE / regulariser / Adam, against
    lp_loss_fwd_bwd_sched -> act_bwd -> conv_wgrad -> adaround_bwd_adam_sched."""
    x, w, b, d = _layer(ops, dev, transposed, N=4, act=act)
    axis = 1 if transposed else 0
    g = torch.Generator().manual_seed(11)
    delta, zp = ops.wq_init_minmax(w, axis, 4)
    alpha0 = ops.adaround_init_alpha(w, delta, axis) + torch.randn(w.shape, generator=g).to(dev)
    sched = ops.new_sched(dev)
    for _ in range(30):                                         # past the warm-up: the regulariser is on
        ops.sched_tick(sched, 100, 0.2, 20, 2)
    wq = ops.adaround_fwd(w, alpha0, delta, zp, axis, 16, True)
    y, fwd_ws = (ops.deconv2d_raw if transposed else ops.conv2d_raw)(x, wq, b, d, want_ws=True)
    tgt_pool = torch.randn(6, *y.shape[1:], generator=g).to(dev)
    table = torch.stack([torch.randperm(6, generator=g)[:4] for _ in range(5)]).to(dev)
    denom = y.numel() // y.shape[1]
    # un-fused chain
    loss_ref, dy = ops.lp_loss_fwd_bwd(y, tgt_pool, 2.0, 1.0 / denom, 2.0 / denom, pick=(table, 3, 1, sched))
    if act:
        dy = ops.act_bwd(y, dy, act, 0.01)
    dw_ref = torch.empty_like(w)
    ops._wgrad(d, transposed, x, dy, dw_ref, (None, 0))
    a_ref, m_ref, v_ref = alpha0.clone(), torch.zeros_like(w), torch.zeros_like(w)
    reg_ref = torch.zeros(1, device=dev)
    ops.adaround_bwd_adam_sched(w, a_ref, delta, zp, dw_ref, m_ref, v_ref, axis, 16, sched, reg_weight=0.01,
                                reg_loss=reg_ref)
    # fused chain
    ws_f = ops._workspace(d, ops.fwd_op(transposed), dev)
    x_slot = ops.conv_x_slot(d, transposed, ws_f)
    ops.stage_mix_sched(x, x, None, 4, 1.0, 0, 1, 0, None, x_slot)
    ws_w = ops._workspace(d, ops.wgrad_op(transposed), dev)
    dy_slot = ops.conv_dy_slot(d, transposed, ws_w)
    assert dy_slot is not None
    loss = torch.zeros(1, device=dev)
    d_pred = torch.empty_like(y)
    ops.lp_loss_stage_sched(y, tgt_pool, table, 3, 1, sched, 2.0, 1.0 / denom, 2.0 / denom, act, 0.01, loss, dy_slot,
                            d_pred=d_pred)
    assert torch.equal(d_pred, dy)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())        # fp32 atomics: order of the partial sums
    a, m, v = alpha0.clone(), torch.zeros_like(w), torch.zeros_like(w)
    reg = torch.zeros(1, device=dev)
    dw = torch.empty_like(w)
    ops.conv_wgrad_adam_sched(d, transposed, x_slot, None, ws_w, w, a, delta, zp, m, v, axis, 16, sched, reg_weight=0.01,
                              reg_loss=reg, dw_out=dw)
    assert torch.equal(dw, dw_ref)
    assert torch.equal(a, a_ref) and torch.equal(m, m_ref) and torch.equal(v, v_ref)
    assert not torch.equal(a, alpha0)
    assert abs(reg.item() - reg_ref.item()) <= 1e-5 * abs(reg_ref.item())


def test_folded_tap_and_simt_layers_have_no_prepared_path(ops, dev):
    x, w, b, d = _layer(ops, dev, False, Cin=3, Cout=32)                       # g_a.0-like: folded taps
    assert ops.packed_weight_bytes(d, False) == 0 and ops.new_packed(d, False, dev) is None
    ws = ops._workspace(d, ops.fwd_op(False), dev)
    assert ops.conv_x_slot(d, False, ws) is None
    x, w, b, d = _layer(ops, dev, False)
    d.engine = ops.ENGINE_SIMT
    assert ops.packed_weight_bytes(d, False) == 0


def _session(dev, fused, host_caches=False):
    from rdo_ptq_b200 import codec, quantization as Q
    from rdo_ptq_b200.quantization import recon
    from rdo_ptq_b200.quantization.session import CalibrationSession
    torch.manual_seed(1005)
    m = codec.ARCHS["mbt2018-mean"](N=64, M=64).eval()
    synth.init_weights(m, gain=1.2)
    m.to(dev)
    q = Q.QuantModel(m, dict(n_bits=4, channel_wise=True, scale_method="max"),
                     dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)).eval()
    cali = synth.calibration_patches(4, 64).to(dev)
    old = recon.FUSED_DEFAULT
    recon.FUSED_DEFAULT = fused
    try:
        s = CalibrationSession(q, cali.cpu() if host_caches == "stream" else cali, batch_size=2, iters=40, warmup=0.1,
                               host_caches=host_caches, n_streams=1)
    finally:
        recon.FUSED_DEFAULT = old
    return s, q


@pytest.mark.parametrize("host_caches", [False, "stream"])
def test_fused_calibration_iteration_is_bit_identical(dev, host_caches):
    """The whole session, fused against un-fused, 8 sweeps (eager + graph replays, past the warm-up): alpha and the Adam
    moments of every unit bit for bit; the fused plan is in use on all 20 units of this codec (conv / transposed conv,
    GDN, and the two folded-tap 3-channel layers)."""
    sa, qa = _session(dev, True, host_caches)
    sb, qb = _session(dev, False, host_caches)
    for _ in range(8):
        sa.sweep()
        sb.sweep()
    torch.cuda.synchronize()
    n_fused = sum(1 for t in sa.trainers.values() if t._fused_plan not in (False, None))
    print(f"fused units: {n_fused} of {len(sa.trainers)}")
    assert n_fused == len(sa.trainers) and all(t._fused_plan in (False, None) for t in sb.trainers.values())
    for n in sa.trainers:
        ta, tb = sa.trainers[n], sb.trainers[n]
        for ma, mb in zip(ta.mods, tb.mods):
            assert torch.equal(ma.weight_quantizer.alpha.data, mb.weight_quantizer.alpha.data), n
        for a, b in zip(ta.exp_avg + ta.exp_avg_sq, tb.exp_avg + tb.exp_avg_sq):
            assert torch.equal(a, b), n
    la, lb = sa.losses(), sb.losses()
    for n in la:
        assert la[n]["rec"] == pytest.approx(lb[n]["rec"], rel=1e-4, abs=1e-9), n
        assert la[n]["round"] == pytest.approx(lb[n]["round"], rel=1e-4, abs=1e-9), n


def test_evaluation_forward_reuses_the_prepared_weights(dev):
    """QuantModule keeps the packed weight operand while nothing it depends on changes: the second forward launches no
    weight kernel, results equal the first forward bit for bit, and a change of alpha rebuilds the operand."""
    from rdo_ptq_b200 import _lib, codec, quantization as Q
    torch.manual_seed(1005)
    m = codec.ARCHS["mbt2018-mean"](N=64, M=64).eval()
    synth.init_weights(m, gain=1.2)
    m.to(dev)
    q = Q.QuantModel(m, dict(n_bits=8, channel_wise=True, scale_method="max"),
                     dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)).eval()
    x = synth.synthetic_image(64, 96).to(dev)
    q.set_quant_state(True, False)
    with torch.no_grad():
        q(x)                                                      # initialises the ranges (general path)
        n0 = _lib.launch_count()
        a = q(x)["x_hat"].clone()                                 # builds the prepared operands
        n1 = _lib.launch_count()
        b = q(x)["x_hat"].clone()                                 # reuses them
        n2 = _lib.launch_count()
    assert torch.equal(a, b)
    mods = [mm for mm in q.modules() if isinstance(mm, Q.QuantModule) and mm.org_weight is not None]
    n_prep = sum(1 for mm in mods if mm.__dict__.get("_prep", {}).get(True, (None, None))[1] is not None)
    assert n_prep >= 14
    assert (n2 - n1) <= (n1 - n0) - n_prep                        # at least one weight kernel per prepared layer is gone
    # against the un-prepared path (gradient enabled => general path): same values up to the two-pass / three-pass split
    q2 = q.model.g_a[2]
    xin = torch.randn(1, 64, 16, 24, device=dev)
    with torch.no_grad():
        y_prep = q2(xin)
    y_gen = q2(xin.requires_grad_(True)).detach()
    assert ((y_prep - y_gen).norm() / y_gen.norm()).item() < 1e-5
    # AdaRound: hard weights are cached, a changed alpha invalidates the cache
    q2.weight_quantizer = Q.AdaRoundQuantizer(q2.weight_quantizer, q2.org_weight.data, round_mode="learned_hard_sigmoid")
    with torch.no_grad():
        y1 = q2(xin.detach()).clone()
        q2.weight_quantizer.alpha.neg_()                  # in-place op on the Parameter: its version counter moves
        y2 = q2(xin.detach()).clone()
        q2.weight_quantizer.alpha.data.neg_()             # write through .data (like a kernel): invisible to the counter
        q2.invalidate_prepared()
        y3 = q2(xin.detach()).clone()
    assert not torch.equal(y1, y2) and torch.equal(y1, y3)
    assert "_prep" in q2.__dict__ and "_prep" not in q2.__getstate__()          # device scratch is not pickled


@pytest.mark.parametrize("arch,kw", [("mbt2018-mean", dict(N=64, M=64)), ("bmshj2018-hyperprior", dict(N=64, M=64)),
                                     ("cheng2020-attn", dict(N=64))])
def test_deferred_activation_quantiser_is_bit_identical(dev, arch, kw):
    """W8A8 evaluation forward with the dynamic activation quantiser of every QuantModule inside an nn.Sequential deferred
    into its consumer's operand staging (ops.DEFER_ACTQ: statistics at the producer, actq_apply_stage at the consumer)
    against the plain form (stats + apply at the producer, NHWC split at the consumer): same codes, so every output
    tensor is bit-identical; fewer launches."""
    from rdo_ptq_b200 import _lib, codec, ops, quantization as Q
    torch.manual_seed(1005)
    m = codec.ARCHS[arch](**kw).eval()
    synth.init_weights(m, gain=1.2)
    m.to(dev)
    q = Q.QuantModel(m, dict(n_bits=8, channel_wise=True, scale_method="max"),
                     dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False),
                     is_cheng=arch.startswith("cheng")).eval()
    x = synth.synthetic_image(128, 192).to(dev)
    q.set_quant_state(True, False)
    with torch.no_grad():
        q(x)
        for mm in q.modules():
            if hasattr(mm, "trained"):
                mm.trained = True
        q.set_quant_state(True, True)
        last = q.model.g_s[-1]
        (last[0] if isinstance(last, torch.nn.Sequential) else last).set_quant_state(True, False)
        q(x)                                                       # prepared weight operands exist from here on
        plain = q(x)
        old, ops.DEFER_ACTQ_MIN_BYTES = ops.DEFER_ACTQ_MIN_BYTES, 0        # defer every eligible layer of this small model
        try:
            n1 = _lib.launch_count()
            with ops.defer_actq():
                deferred = q(x)
            n2 = _lib.launch_count()
            q(x)
            n3 = _lib.launch_count()
        finally:
            ops.DEFER_ACTQ_MIN_BYTES = old
    assert n2 - n1 != n3 - n2                                              # the deferred form really ran
    linked = sum(1 for mm in q.modules() if isinstance(mm, Q.QuantModule) and "_defer_to" in mm.__dict__)
    assert linked >= (14 if not arch.startswith("cheng") else 8)
    assert torch.equal(plain["x_hat"], deferred["x_hat"])
    assert torch.equal(plain["likelihoods"]["y"], deferred["likelihoods"]["y"])
    assert torch.equal(plain["likelihoods"]["z"], deferred["likelihoods"]["z"])
    assert getattr(deferred["x_hat"], "_b200_actq", None) is None          # nothing pending leaves the model


@pytest.mark.parametrize("shape", [
    # (N, Cin, H, W, Cout, k, stride): 128 equal items (the g_a.2 class), a handful of pixel tiles (h_a.2 class), a
    # ragged shape, and the integer-weight form
    (8, 192, 64, 64, 192, 5, 2), (8, 192, 16, 16, 192, 5, 2), (2, 96, 30, 44, 160, 3, 1), (8, 320, 16, 16, 192, 3, 1)])
def test_stream_k_schedule_matches_whole_item_schedule(ops, dev, shape):
    """Stream-K scheduling of the conv engine (K ranges of the items laid end to end, partial accumulators summed by the
    item's owner) against whole-item scheduling: same MMAs, the K blocks of an output element are summed in a different
    grouping, so results agree to fp32 rounding (1e-6 relative), for the three-pass and the integer two-pass form."""
    from rdo_ptq_b200 import _lib
    N, Cin, H, W, Cout, k, st = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    d = ops.conv_desc(x.shape, w.shape, st, k // 2, act=2, slope=0.01)
    delta, zp = ops.wq_init_minmax(w, 0, 8)
    n_int = ops.wq_int_weights(w, delta, zp, 0, 256)
    outs = {}
    try:
        for mode in (0, 2):
            assert _lib.lib().b200lic_set_option(b"streamk", mode) == 0
            y = ops.conv2d_raw(x, w, b, d)
            y_int = ops.conv_wq(x, n_int, delta.reshape(-1).contiguous(), b, stride=st, padding=k // 2, act=2, slope=0.01)
            torch.cuda.synchronize()
            outs[mode] = (y, y_int)
    finally:
        _lib.lib().b200lic_set_option(b"streamk", 1)
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), st, k // 2), 0.01)
    wq = (n_int * delta).double()
    ref_int = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), wq, b.double(), st, k // 2), 0.01)
    errs = {}
    for mode in (0, 2):
        errs[mode] = (((outs[mode][0].double() - ref).norm() / ref.norm()).item(),
                      ((outs[mode][1].double() - ref_int).norm() / ref_int.norm()).item())
    between = [((a - c).norm() / a.norm()).item() for a, c in zip(outs[0], outs[2])]
    print(f"{shape}: vs fp64 whole-item {errs[0]}, stream-K {errs[2]}; between the schedules {between}")
    for mode in (0, 2):
        assert errs[mode][0] < 2e-5 and errs[mode][1] < 2e-5               # both well inside the 1e-4 per-layer bar
    assert max(between) < 4e-5


@pytest.mark.parametrize("shape", [
    # (N, Cin, H, W, Cout, k, stride, transposed): the g_a.2 class (many pixel tiles), an odd tile count (the last pair has
    # one tile only), a ragged shape, a transposed conv with ragged phases, a 32-channel tile, few tiles x wide Cout
    (8, 192, 64, 64, 192, 5, 2, False), (1, 64, 24, 16, 96, 3, 1, False), (2, 96, 30, 44, 160, 3, 1, False),
    (1, 64, 9, 13, 96, 5, 2, True), (1, 32, 16, 16, 32, 3, 1, False), (8, 480, 16, 16, 640, 3, 1, False),
    (8, 192, 32, 32, 192, 5, 2, True)])
def test_cta_pair_form_matches_the_single_cta_form(ops, dev, shape):
    """The CTA-pair form of the conv engine (clusters of two CTAs, tcgen05 cta_group::2: M = 256 per MMA, each CTA staging
    its own pixel tile and half of the weight tile) against the single-CTA form: every output element is accumulated from
    the same products in the same K order, so whole-item schedules agree bit for bit; with stream-K forced on both sides
    (the K ranges are cut differently) to fp32 rounding.  Three-pass and integer two-pass forms, both against fp64."""
    from rdo_ptq_b200 import _lib
    N, Cin, H, W, Cout, k, st, tr = shape
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ax = 1 if tr else 0
    delta, zp = ops.wq_init_minmax(w, ax, 8)
    n_int = ops.wq_int_weights(w, delta, zp, ax, 256)
    F = torch.nn.functional
    if tr:
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), st, k // 2, st - 1)
        ref_int = F.conv_transpose2d(x.double(), (n_int * delta).double(), b.double(), st, k // 2, st - 1)
    else:
        ref = F.conv2d(x.double(), w.double(), b.double(), st, k // 2)
        ref_int = F.conv2d(x.double(), (n_int * delta).double(), b.double(), st, k // 2)
    outs = {}
    try:
        for sk in (0, 2):
            for pair in (0, 2):
                assert _lib.lib().b200lic_set_option(b"streamk", sk) == 0
                assert _lib.lib().b200lic_set_option(b"pair", pair) == 0
                d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0)
                y = ops.deconv2d_raw(x, w, b, d) if tr else ops.conv2d_raw(x, w, b, d)
                y_int = ops.conv_wq(x, n_int, delta.reshape(-1).contiguous(), b, stride=st, padding=k // 2,
                                    output_padding=st - 1 if tr else 0, transposed=tr)
                torch.cuda.synchronize()
                outs[(sk, pair)] = (y, y_int)
    finally:
        _lib.lib().b200lic_set_option(b"streamk", 1)
        _lib.lib().b200lic_set_option(b"pair", 1)
    for key, (y, y_int) in outs.items():
        assert ((y.double() - ref).norm() / ref.norm()).item() < 2e-5, key
        if y_int is not None:
            assert ((y_int.double() - ref_int).norm() / ref_int.norm()).item() < 2e-5, key
    assert torch.equal(outs[(0, 0)][0], outs[(0, 2)][0])                   # whole items: same accumulation chains
    if outs[(0, 0)][1] is not None:
        assert torch.equal(outs[(0, 0)][1], outs[(0, 2)][1])
    assert ((outs[(2, 0)][0] - outs[(2, 2)][0]).norm() / outs[(2, 0)][0].norm()).item() < 4e-5


@pytest.mark.parametrize("shape", [
    # (N, Cin, H, W, Cout, k, stride, transposed) with an ODD number of pixel tiles at the widest channel tile: the peer
    # CTA of the last pair has nothing to write (3 tiles; 5 tiles; 9 + ragged phases of a transposed conv)
    (1, 64, 24, 16, 96, 3, 1, False), (1, 96, 40, 16, 128, 3, 1, False), (1, 64, 17, 24, 64, 5, 2, True)])
def test_cta_pair_form_with_an_odd_tile_count(ops, dev, shape, monkeypatch):
    """The widest output-channel tile is forced (B200LIC_TC_BN, an experiments knob: the cost model would give these few-
    pixel shapes 16-channel tiles, which the pair form does not take), so that the pair form meets an odd pixel-tile count:
    the last pair's second CTA loads a duplicate tile, joins the M = 256 MMAs and stores nothing."""
    from rdo_ptq_b200 import _lib
    monkeypatch.setenv("B200LIC_TC_BN", "256")
    N, Cin, H, W, Cout, k, st, tr = shape
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    F = torch.nn.functional
    ref = (F.conv_transpose2d(x.double(), w.double(), b.double(), st, k // 2, st - 1) if tr
           else F.conv2d(x.double(), w.double(), b.double(), st, k // 2))
    outs = {}
    try:
        for pair in (0, 2):
            assert _lib.lib().b200lic_set_option(b"pair", pair) == 0
            d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0)
            info = ops.conv_plan_info(d, tr)
            assert info["pair"] == (1 if pair else 0) and info["BN"] == Cout
            if not tr:
                assert info["m_tiles"] % 2 == 1
            outs[pair] = ops.deconv2d_raw(x, w, b, d) if tr else ops.conv2d_raw(x, w, b, d)
            torch.cuda.synchronize()
    finally:
        _lib.lib().b200lic_set_option(b"pair", 1)
    assert ((outs[2].double() - ref).norm() / ref.norm()).item() < 2e-5
    assert torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("case", [
    # N, Cin, H, W, Cout, k, stride, transposed, act: generic engine (pairs / single CTA), ragged tiles, padded channel
    # count, transposed conv (strided phase stores), 1x1 short-K engine, folded 3 -> N conv, stream-K owner
    (2, 96, 30, 44, 160, 3, 1, False, 2), (1, 64, 24, 16, 40, 3, 1, False, 0), (1, 64, 9, 13, 96, 5, 2, True, 1),
    (2, 192, 16, 16, 192, 1, 1, False, 2), (1, 3, 64, 96, 128, 5, 2, False, 0), (8, 192, 32, 32, 192, 5, 2, False, 0)])
@pytest.mark.parametrize("pair", [0, 2])
def test_activation_statistics_from_the_conv_epilogue(ops, dev, case, pair):
    """b200lic_conv_stats_once: the per-channel (min, max) keys the convolution's epilogue merges from its registers equal
    the keys of b200lic_actq_stats over the stored output, bit for bit (min / max are exact), on every path that honours
    the request; an engine that does not (SIMT) leaves it pending."""
    from rdo_ptq_b200 import _lib
    N, Cin, H, W, Cout, k, st, tr, act = case
    g = torch.Generator().manual_seed(21 + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ax = 1 if tr else 0
    delta, zp = ops.wq_init_minmax(w, ax, 8)
    n_int = ops.wq_int_weights(w, delta, zp, ax, 256)
    try:
        assert _lib.lib().b200lic_set_option(b"pair", pair) == 0
        for sk in (1, 2):
            assert _lib.lib().b200lic_set_option(b"streamk", sk) == 0
            d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0, act=act, slope=0.01)
            keys = ops.conv_stats_arm(Cout, dev)
            y = ops.deconv2d_raw(x, w, b, d) if tr else ops.conv2d_raw(x, w, b, d)
            folded_deconv = tr and not ops.conv_stats_taken() if tr else False
            if not tr:
                assert ops.conv_stats_taken()
            if not folded_deconv:
                assert torch.equal(keys, ops.act_quant_stats(y)), (sk, "three-pass")
            keys = ops.conv_stats_arm(Cout, dev)
            y_int = ops.conv_wq(x, n_int, delta.reshape(-1).contiguous(), b, stride=st, padding=k // 2,
                                output_padding=st - 1 if tr else 0, transposed=tr, act=act, slope=0.01)
            taken = ops.conv_stats_taken()
            if y_int is not None and taken:
                assert torch.equal(keys, ops.act_quant_stats(y_int)), (sk, "integer")
        d = ops.conv_desc(x.shape, w.shape, st, k // 2, tr, st - 1 if tr else 0, engine=ops.ENGINE_SIMT)
        keys = ops.conv_stats_arm(Cout, dev)
        (ops.deconv2d_raw if tr else ops.conv2d_raw)(x, w, b, d)
        assert not ops.conv_stats_taken()                      # the exact-fp32 engine has no fused statistics
        assert _lib.lib().b200lic_conv_stats_pending() == 0    # ... and the request is disarmed
    finally:
        _lib.lib().b200lic_set_option(b"streamk", 1)
        _lib.lib().b200lic_set_option(b"pair", 1)


@pytest.mark.parametrize("shape", [(1, 192, 64, 96), (3, 80, 30, 22), (2, 256, 40, 36)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("deferred", [False, True])
def test_activation_statistics_from_the_gdn_epilogue(ops, dev, shape, inverse, deferred):
    """The one-kernel GDN honours b200lic_conv_stats_once too: keys of its output == b200lic_actq_stats(y), bit for bit
    (ragged pixel tails, padded channel counts, with and without the deferred input quantiser)."""
    N, Cc, H, W = shape
    g = torch.Generator().manual_seed(5 + Cc)
    x = (torch.randn(N, Cc, H, W, generator=g) * 2).to(dev)
    gam = (torch.rand(Cc, Cc, generator=g) * 0.02 + 0.1 * torch.eye(Cc)).to(dev)
    bet = (1 + torch.rand(Cc, generator=g)).to(dev)
    d = ops.gdn_desc(x.shape, inverse)
    packed = ops.pack_weights(gam.view(Cc, Cc, 1, 1), d, False)
    pend = (ops.act_quant_stats(x), 8) if deferred else None
    keys = ops.conv_stats_arm(Cc, dev)
    y = ops.gdn_fwd_fused(x, packed, bet, inverse, pending=pend)
    assert ops.conv_stats_taken()
    assert torch.equal(keys, ops.act_quant_stats(y))


@pytest.mark.parametrize("shape", [(1, 64, 24, 40, 128, 12), (2, 96, 17, 23, 192, 12), (1, 32, 16, 16, 64, 13)])
def test_masked_context_conv_contracts_the_live_taps_only(ops, dev, shape):
    """b200lic_conv_desc::k_taps: the 5x5 context convolution behind compressai's causal mask (mask 'A': the first 12
    taps in raster order, 'B': 13) on a weight whose masked taps are zero -- the truncated K loop adds exact zeros less,
    so the output is bit-identical to the dense contraction; on a weight with something behind the mask it equals the
    convolution with that part removed (the caller's assertion is what makes the two the same)."""
    N, Cin, H, W, Cout, L = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 5, 5, generator=g) * 0.05).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    w_masked = w.clone()
    w_masked.view(Cout, Cin, 25)[:, :, L:] = 0
    d = ops.conv_desc(x.shape, w.shape, 1, 2)
    dk = ops.conv_desc(x.shape, w.shape, 1, 2, k_taps=L)
    dense = ops.conv_fwd_packed(x, ops.pack_weights(w_masked, d, False), d, False, bias=b)
    live = ops.conv_fwd_packed(x, ops.pack_weights(w_masked, dk, False), dk, False, bias=b)
    assert torch.equal(dense, live)
    ref = torch.nn.functional.conv2d(x.double(), w_masked.double(), b.double(), 1, 2)
    assert ((live.double() - ref).norm() / ref.norm()).item() < 2e-5
    trunc = ops.conv_fwd_packed(x, ops.pack_weights(w, dk, False), dk, False, bias=b)     # something behind the mask
    assert torch.equal(trunc, live)


def test_quantised_cheng2020_context_model_runs_on_the_live_taps(dev):
    """The wrapped MaskedConv2d of Cheng2020 (TO quant_model.py:45-48 wraps it as a dense conv): the evaluation forward
    contracts 12 of its 25 taps when the quantised weight is zero behind the mask, all 25 after a masked tap has been
    set (AdaRound may un-mask one); outputs equal the dense form bit for bit."""
    from rdo_ptq_b200 import codec, quantization as Q, synth
    from rdo_ptq_b200.quantization import quant_layer as QL
    torch.manual_seed(1005)
    m = codec.ARCHS["cheng2020-attn"](N=32).eval()
    synth.init_weights(m, gain=0.6)
    m.to(dev)
    x = synth.synthetic_image(64, 128).to(dev)
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    with torch.no_grad():
        m(x)                                              # bakes the mask into weight.data
        qnn = Q.QuantModel(m, wq, aq, is_cheng=True).eval()
        qnn.set_quant_state(True, False)
        ctx = qnn.model.context_prediction
        assert isinstance(ctx, Q.QuantModule) and ctx.mask_live_taps == 12
        qnn(x)                                            # first quantised forward initialises the weight quantisers
        out = qnn(x)                                      # ... the second runs on prepared operands
        assert ctx.last_k_taps == 12
        QL.MASKED_TAPS = False
        try:
            for mod in qnn.modules():
                if isinstance(mod, Q.QuantModule):
                    mod.invalidate_prepared()
            dense = qnn(x)
            assert ctx.last_k_taps == 0
        finally:
            QL.MASKED_TAPS = True
        assert torch.equal(out["x_hat"], dense["x_hat"])
        assert torch.equal(out["likelihoods"]["y"], dense["likelihoods"]["y"])
        # a masked tap that holds something (what a trained alpha can do): dense contraction again
        ctx.weight.data[:, :, 4, 4] = ctx.weight.data[:, :, 0, 0]
        ctx.invalidate_prepared()
        qnn(x)
        assert ctx.last_k_taps == 0


@pytest.mark.parametrize("case", [
    # N, Cin, H, W, Cout, transposed, act, integer
    (2, 75, 24, 40, 192, False, 2, False),      # folded 3 -> N analysis conv (K = 75 -> 96)
    (1, 192, 33, 21, 75, True, 0, False),       # folded N -> 3 synthesis layer: col = x . W'' (ragged pixel tail)
    (2, 192, 16, 16, 192, False, 0, False),     # GDN norm GEMM of the calibration
    (1, 64, 20, 12, 40, False, 1, True),        # integer two-pass form, Cout padded to 48
    (3, 256, 9, 7, 256, False, 0, False),       # widest eligible shape
])
def test_short_k_1x1_engine_matches_the_generic_engine(ops, dev, case):
    """gemm1x1_tc.cu (weights resident in shared memory, register epilogue) against conv_tc2.cu on the same staged
    operands: same products in the same order, so the outputs are bit-identical; and both against fp64."""
    from rdo_ptq_b200 import _lib
    N, Cin, H, W, Cout, tr, act, integer = case
    g = torch.Generator().manual_seed(5 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    wshape = (Cin, Cout, 1, 1) if tr else (Cout, Cin, 1, 1)
    w = (torch.randn(wshape, generator=g) * 0.1)
    if integer:
        w = torch.randint(-128, 128, wshape, generator=g).float()
    w = w.to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    scale = (torch.rand(Cout, generator=g) * 0.01 + 0.001).to(dev) if integer else None
    d = ops.conv_desc(x.shape, w.shape, 1, 0, tr, 0, act=act, slope=0.01)
    packed = ops.pack_weights(w, d, tr)
    outs = []
    for mode in (1, 0):
        assert _lib.lib().b200lic_set_option(b"gemm1x1", mode) == 0
        try:
            n0 = _lib.launch_count()
            outs.append(ops.conv_fwd_packed(x, packed, d, tr, bias=b, w_scale=scale))
        finally:
            _lib.lib().b200lic_set_option(b"gemm1x1", 1)
    assert torch.equal(outs[0], outs[1])
    w2 = w.view(Cin, Cout).t() if tr else w.view(Cout, Cin)
    ref = torch.einsum("ok,nkhw->nohw", w2.double(), x.double())
    if integer:
        ref = ref * scale.double().view(1, -1, 1, 1)
    ref = ref + b.double().view(1, -1, 1, 1)
    ref = torch.relu(ref) if act == 1 else (torch.where(ref > 0, ref, ref * 0.01) if act == 2 else ref)
    assert (outs[0].double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()
