"""CPU: self-checks of the compressai restatement (oracle.codec) against closed forms -- the reference holds no
fixture for these pieces ("parity unpinned")."""
import math

import pytest

import numpy as np
import torch
from scipy import special

from oracle import codec, evalpath, quant_wrap, calib
from rdo_ptq_b200 import synth


def test_gaussian_likelihood_is_a_pmf_and_matches_scipy():
    gc = codec.GaussianConditional(None).eval()
    ks = torch.arange(-60, 61, dtype=torch.float32).view(1, 1, -1, 1)
    for sigma, mu in ((0.11, 0.3), (1.0, -0.2), (4.0, 0.45)):
        y = ks + mu
        _, lik = gc(y, torch.full_like(y, sigma), means=torch.full_like(y, mu))
        assert abs(lik.sum().item() - 1.0) < 1e-4
        k = ks.flatten().double().numpy()
        ref = 0.5 * special.erfc(-(0.5 - np.abs(k)) / sigma / math.sqrt(2)) - 0.5 * special.erfc(-(-0.5 - np.abs(k)) / sigma / math.sqrt(2))
        assert np.allclose(lik.flatten().double().numpy(), np.maximum(ref, 1e-9), atol=2e-7)


def test_scale_lower_bound_and_rounding():
    gc = codec.GaussianConditional(None).eval()
    y = torch.tensor([[[[0.5, 1.5, 2.5, -0.5]]]])
    yh, lik = gc(y, torch.full_like(y, 1e-3))
    assert yh.flatten().tolist() == [0.0, 2.0, 2.0, -0.0]
    _, lik_b = gc(y, torch.full_like(y, 0.11))
    assert torch.equal(lik, lik_b)


def test_factorized_prior_is_monotone_and_sums_to_one(golden_c):
    eb = codec.EntropyBottleneck(5).eval()
    eb.load_state_dict(golden_c["eb"]["state"])
    v = torch.linspace(-40, 40, 801).view(1, 1, -1).repeat(5, 1, 1)
    c = eb._logits_cumulative(v)
    assert (c[:, :, 1:] >= c[:, :, :-1]).all()
    med = eb._get_medians().detach()
    ks = torch.arange(-200, 201, dtype=torch.float32).view(1, 1, -1) + med
    lik = eb._likelihood(ks)
    assert torch.allclose(lik.sum(-1), torch.ones(5, 1), atol=1e-3)


def test_entropy_bottleneck_eval_rounds_around_median(golden_c):
    g = golden_c["eb"]
    eb = codec.EntropyBottleneck(5).eval()
    eb.load_state_dict(g["state"])
    zh, lik = eb(g["z"])
    med = eb._get_medians().detach().view(1, 5, 1, 1)
    assert torch.equal(zh, torch.round(g["z"] - med) + med)
    assert torch.equal(zh, g["z_hat"]) and torch.allclose(lik, g["lik"], rtol=1e-6, atol=1e-9)
    assert (lik >= 1e-9).all() and (lik <= 1).all()


def test_gdn_igdn_round_trip_on_diagonal_gamma():
    gdn, igdn = codec.GDN(4), codec.GDN(4, inverse=True)
    x = torch.randn(2, 4, 5, 5)
    y = gdn(x)                                   # y = x / sqrt(1 + 0.1 x^2)  =>  x = y / sqrt(1 - 0.1 y^2)
    assert torch.allclose(y, x / torch.sqrt(1 + 0.1 * x * x), atol=1e-6)
    assert torch.allclose(igdn(x), x * torch.sqrt(1 + 0.1 * x * x), atol=1e-5)
    back = y / torch.sqrt(1 - 0.1 * y * y)
    assert torch.allclose(back, x, atol=1e-4)


def test_masked_conv_is_causal():
    m = codec.MaskedConv2d(2, 3, kernel_size=5, padding=2, bias=False)
    x = torch.zeros(1, 2, 9, 9)
    x[0, :, 4, 4] = 1.0
    y = m(x)
    assert (m.mask.sum(dim=(2, 3)) == 12).all()
    assert y[0, :, :4].abs().sum() == 0 and y[0, :, 4, :5].abs().sum() == 0     # nothing above / left / at centre
    assert y[0, :, 5:7].abs().sum() > 0


def test_pad_crop_round_trip_and_bpp():
    x = torch.rand(1, 3, 70, 100)
    xp = evalpath.pad(x, 64)
    assert xp.shape == (1, 3, 128, 128)
    assert torch.equal(evalpath.crop(xp, (70, 100)), x)
    out = {"x_hat": xp, "likelihoods": {"y": torch.full((1, 2, 8, 8), 0.5)}}
    assert abs(evalpath.compute_bpp(out) - 128 / (128 * 128)) < 1e-9


def test_golden_models_are_reproducible(golden_c):
    for arch in ("mbt2018-mean", "bmshj2018-hyperprior", "cheng2020-attn"):
        g = golden_c[f"model/{arch}"]
        m = codec.ARCHS[arch](**g["kw"]).eval()
        m.load_state_dict(g["state"])
        with torch.no_grad():
            out = m(g["x"])
        assert torch.allclose(out["x_hat"], g["x_hat"], rtol=1e-4, atol=1e-5)
        assert abs(evalpath.compute_bpp(out) - g["bpp"]) < 1e-3


def test_quant_model_rewrite_rules():
    m = codec.Cheng2020Attention(N=12).eval()
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    q = quant_wrap.QuantModel(m, wq, aq)
    assert isinstance(q.model.g_a[0], quant_wrap.QuantRBWS) and isinstance(q.model.g_s[2], quant_wrap.QuantRBU)
    assert isinstance(q.model.h_a[1], quant_wrap.StraightThrough)            # LeakyReLU absorbed
    assert isinstance(q.model.h_a[0].activation_function, torch.nn.LeakyReLU)
    ps = q.model.g_s[9][1]
    assert isinstance(ps, quant_wrap.QuantModule) and ps.is_ps                # Q4: wrapped PixelShuffle + LeakyReLU
    assert isinstance(q.model.context_prediction, quant_wrap.QuantModule)     # Q5: mask dropped by the wrap
    n_blocks = sum(isinstance(x, quant_wrap.BaseQuantBlock) for x in q.model.modules())
    assert n_blocks == 13


def test_oracle_calibration_reduces_reconstruction_error():
    torch.manual_seed(0)
    m = codec.MeanScaleHyperprior(N=8, M=12).eval()
    synth.init_weights(m, gain=1.2)
    wq = dict(n_bits=4, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    q = quant_wrap.QuantModel(m, wq, aq)
    q.eval()
    cali = synth.calibration_patches(4, 64)
    q.set_quant_state(True, False)
    with torch.no_grad():
        q(cali[:2])
    layer = q.model.g_a[2]
    losses = calib.reconstruct(q, layer, 0, "2", cali, batch_size=2, iters=100, weight=0.0, warmup=0.2, input_prob=1.0)
    assert layer.trained and not layer.weight_quantizer.soft_targets
    assert sum(losses[-10:]) < sum(losses[:10])


def test_oracle_rate_gradients_match_finite_differences():
    """The autograd the GPU backward kernels are checked against: d(-log2 lik)/d(scale) of the Gaussian conditional and
    d(-log2 lik)/dz of the factorised prior under straight-through rounding, against central differences in fp64."""
    torch.manual_seed(0)
    gc = codec.GaussianConditional(None).eval().double()
    y = torch.tensor([[[[0.3, -2.2, 4.1, 0.0]]]], dtype=torch.float64)
    mu = torch.tensor([[[[0.1, 0.4, -0.3, 0.2]]]], dtype=torch.float64)
    sc = torch.tensor([[[[0.5, 1.3, 2.0, 0.3]]]], dtype=torch.float64, requires_grad=True)

    def bits(s):
        return (-torch.log2(gc(y, s, means=mu)[1])).sum()
    bits(sc).backward()
    eps = 1e-6
    for i in range(4):
        d = torch.zeros_like(sc)
        d[..., i] = eps
        fd = (bits(sc.detach() + d) - bits(sc.detach() - d)) / (2 * eps)
        assert abs(fd.item() - sc.grad[..., i].item()) < 1e-6 * max(1.0, abs(fd.item()))
    eb = codec.EntropyBottleneck(3).eval().double()
    with torch.no_grad():
        for i in range(4):
            getattr(eb, f"_factor{i}").normal_(0, 0.5)
    eb.ste_round = True
    z = (torch.randn(1, 3, 2, 2, dtype=torch.float64) * 3).requires_grad_(True)
    (-torch.log2(eb(z)[1])).sum().backward()
    # straight-through rounding: the gradient w.r.t. z equals the derivative of -log2 lik(v) at the SYMBOL v = z_hat
    zh = eb(z)[0].detach()

    def bits_at(v):
        vv = v.transpose(0, 1).reshape(3, 1, -1)
        return (-torch.log2(eb._likelihood(vv))).sum()
    for idx in [(0, 0, 0, 0), (0, 1, 1, 0), (0, 2, 1, 1)]:
        d = torch.zeros_like(zh)
        d[idx] = eps
        fd = (bits_at(zh + d) - bits_at(zh - d)) / (2 * eps)
        assert abs(fd.item() - z.grad[idx].item()) < 1e-5 * max(1.0, abs(fd.item()))


def _ms_ssim_direct_f64(x, y, data_range=1.0):
    """Independent evaluation of the published MS-SSIM definition in float64: the full 11x11 window (outer product of
    the 1-D Gaussian) applied through unfold, explicit 2x2 pooling loops -- no code shared with oracle/msssim.py."""
    from oracle.msssim import WEIGHTS
    x, y = x.double(), y.double()
    c = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-(c ** 2) / (2 * 1.5 ** 2))
    g = g / g.sum()
    w2 = torch.outer(g, g).reshape(1, 121, 1)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2

    def filt(t):                                    # [N,C,H,W] -> [N,C,(H-10)*(W-10)]
        n, ch, h, w = t.shape
        cols = torch.nn.functional.unfold(t.reshape(n * ch, 1, h, w), 11)       # [n*ch, 121, L]
        return (cols * w2).sum(1).reshape(n, ch, -1)

    def pool(t):
        n, ch, h, w = t.shape
        ph, pw = h % 2, w % 2
        tp = torch.zeros(n, ch, h + 2 * ph, w + 2 * pw, dtype=t.dtype)
        tp[:, :, ph:ph + h, pw:pw + w] = t
        ho, wo = (h + 2 * ph - 2) // 2 + 1, (w + 2 * pw - 2) // 2 + 1
        tp = tp[:, :, :2 * ho, :2 * wo]
        return 0.25 * (tp[:, :, 0::2, 0::2] + tp[:, :, 0::2, 1::2] + tp[:, :, 1::2, 0::2] + tp[:, :, 1::2, 1::2])

    vals = []
    for lvl in range(5):
        m1, m2 = filt(x), filt(y)
        s1, s2, s12 = filt(x * x) - m1 * m1, filt(y * y) - m2 * m2, filt(x * y) - m1 * m2
        cs = (2 * s12 + c2) / (s1 + s2 + c2)
        ss = (2 * m1 * m2 + c1) / (m1 * m1 + m2 * m2 + c1) * cs
        if lvl < 4:
            vals.append(cs.mean(-1).clamp(min=0))
            x, y = pool(x), pool(y)
        else:
            vals.append(ss.mean(-1).clamp(min=0))
    wts = torch.tensor(WEIGHTS, dtype=torch.float64).view(-1, 1, 1)
    return torch.prod(torch.stack(vals) ** wts, dim=0).mean().item()


def test_ms_ssim_oracle_closed_forms_and_direct_evaluation():
    """oracle/msssim.py (restatement of pytorch_msssim 1.0.0, dependency absent): identical images give exactly 1, more
    noise gives less, the value agrees with an independent float64 evaluation of the definition (odd and even sides:
    both pooling paddings), and the pooling matches avg_pool2d's zero-padded, pad-counted mean."""
    from oracle import msssim
    g = torch.Generator().manual_seed(3)
    x = synth.synthetic_images(1, 177, 200)[0]
    assert abs(msssim.ms_ssim(x, x).item() - 1.0) < 1e-6
    prev = 1.0
    for sigma in (0.01, 0.05, 0.2):
        y = (x + sigma * torch.randn(x.shape, generator=g)).clamp(0, 1)
        v = msssim.ms_ssim(x, y).item()
        assert 0.0 < v < prev
        assert abs(v - _ms_ssim_direct_f64(x, y)) < 2e-5
        prev = v
    x2 = torch.rand(2, 3, 192, 161, generator=g)
    y2 = (x2 + 0.1 * torch.randn(x2.shape, generator=g)).clamp(0, 1)
    assert abs(msssim.ms_ssim(x2, y2).item() - _ms_ssim_direct_f64(x2, y2)) < 2e-5
    per = msssim.ms_ssim(x2, y2, size_average=False)
    assert per.shape == (2,) and abs(per.mean().item() - msssim.ms_ssim(x2, y2).item()) < 1e-6
    with pytest.raises(ValueError):
        msssim.ms_ssim(x2[:, :, :160], y2[:, :, :160])
    t = torch.arange(15.0).reshape(1, 1, 3, 5)
    d = msssim.downsample(t)
    assert d.shape == (1, 1, 2, 3) and d[0, 0, 0, 0].item() == 0.0 and d[0, 0, 1, 1].item() == (6 + 7 + 11 + 12) / 4
