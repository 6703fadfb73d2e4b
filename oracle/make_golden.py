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


def main():
    os.makedirs(OUT, exist_ok=True)
    q = quantizer_vectors()
    torch.save(q, os.path.join(OUT, "quantizer_ref.pt"))
    print(f"quantizer_ref.pt: {len(q)} cases pinned bit-exactly against /root/reference")
    c = codec_vectors()
    torch.save(c, os.path.join(OUT, "codec_oracle.pt"))
    print(f"codec_oracle.pt: {len(c)} cases (oracle-only, unpinned)")


if __name__ == "__main__":
    main()
