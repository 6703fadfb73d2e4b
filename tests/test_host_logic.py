"""CPU: host-side logic of the product (graph rewrite, schedules, sharding) incl. the N>1 path on gloo, world_size 2."""
import os
import socket

import torch
import torch.multiprocessing as mp

from oracle import quant_wrap as owrap, codec as ocodec


def test_product_graph_rewrite_matches_oracle():
    from rdo_ptq_b200 import codec, quantization as Q
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    for arch, kw in (("mbt2018-mean", dict(N=8, M=12)), ("bmshj2018-hyperprior", dict(N=8, M=12)),
                     ("cheng2020-attn", dict(N=12))):
        p = Q.QuantModel(codec.ARCHS[arch](**kw), wq, aq)
        o = owrap.QuantModel(ocodec.ARCHS[arch](**kw), wq, aq)
        ps = [(n, type(m).__name__) for n, m in p.named_modules()
              if isinstance(m, (Q.QuantModule, Q.BaseQuantBlock, Q.StraightThrough))]
        os_ = [(n, type(m).__name__) for n, m in o.named_modules()
               if isinstance(m, (owrap.QuantModule, owrap.BaseQuantBlock, owrap.StraightThrough))]
        assert ps == os_ and len(ps) > 10, arch
        pa = [type(m.activation_function).__name__ for m in p.modules() if isinstance(m, Q.QuantModule)]
        oa = [type(m.activation_function).__name__ for m in o.modules() if isinstance(m, owrap.QuantModule)]
        assert pa == oa
        assert sorted(p.state_dict().keys()) == sorted(o.state_dict().keys())
    from rdo_ptq_b200.quantization.session import reconstruction_units
    p = Q.QuantModel(codec.ARCHS["mbt2018-mean"](N=8, M=12), wq, aq)
    assert len(reconstruction_units(p)) == 20


def test_lu_graph_rewrite_and_schedule():
    from rdo_ptq_b200 import codec, quant_int as LU, quantization as Q
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=True)
    m = LU.QuantModel(codec.ScaleHyperprior(8, 12), wq, aq)
    assert sum(isinstance(x, LU.QuantModule) for x in m.modules()) == 14
    assert sum(isinstance(x, codec.GDN) for x in m.modules()) == 6            # GDN stays fp32
    c = LU.QuantCodingModel(codec.ScaleHyperprior(8, 12), wq, aq)
    assert sum(isinstance(x, LU.QuantModule) for x in c.modules()) == 6       # h_a + h_s only
    d = Q.LinearTempDecay(20000, rel_start_decay=0.2, start_b=20, end_b=2)
    assert d(0) == 20 and abs(d(12000) - 11.0) < 1e-9 and d(20000) == 2
    from rdo_ptq_b200 import evaluate as E
    x = torch.rand(1, 3, 70, 100)
    assert torch.equal(E.crop(E.pad(x, 64), (70, 100)), x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from rdo_ptq_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = D.shard_indices(7)
    grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), float(10 * (rank + 1)))]
    D.allreduce_flat_(grads, average=True)
    psnr, bpp, cnt = D.reduce_metrics(30.0 * len(idx), 0.5 * len(idx), len(idx))
    q.put((rank, idx, grads[0][0, 0].item(), grads[1][0].item(), psnr, bpp, cnt))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]               # round-robin image shards
    for r in res:
        assert r[2] == 1.5 and r[3] == 15.0                                   # mean of the two ranks' gradients
        assert abs(r[4] - 30.0) < 1e-12 and abs(r[5] - 0.5) < 1e-12 and r[6] == 7


def test_entry_point_flags_and_walk_order(monkeypatch):
    """main2.py drop-in: the reference's flag defaults (main2.py:20-60) and the depth-first recon_model walk
    (main2.py:227-250) visiting the units in the oracle's order, layers -> layer_reconstruction, blocks ->
    block_reconstruction, with the walk position as unit_id; quantize.py drop-in: its defaults (quantize.py:27-48)."""
    from oracle import calib as ocalib
    from rdo_ptq_b200 import codec, main2, quantize as lu_entry, quantization as Q
    a = main2.parse_args([])
    assert (a.seed, a.batch_size, a.n_bits_w, a.n_bits_a, a.iters_w, a.num_samples) == (1005, 4, 8, 8, 20000, 12)
    assert (a.weight, a.b_start, a.b_end, a.warmup, a.input_prob, a.lr, a.init, a.task_loss) == \
        (0.01, 20, 2, 0.2, 0.5, 4e-5, "max", 2.0)
    assert not a.channel_wise and not a.act_quant and not a.disable_8bit_head_stem
    assert main2.parse_args(["--task_loss", "rd"]).task_loss == "rd"
    b = lu_entry.parse_args([])
    assert (b.seed, b.n_bits_w, b.n_bits_a, b.type, b.init) == (1005, 8, 16, "INT8", "max")
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    for arch, kw in (("mbt2018-mean", dict(N=8, M=12)), ("cheng2020-attn", dict(N=12))):
        p = Q.QuantModel(codec.ARCHS[arch](**kw), wq, aq)
        o = owrap.QuantModel(ocodec.ARCHS[arch](**kw), wq, aq)
        seen, oseen = [], []
        monkeypatch.setattr(main2, "layer_reconstruction",
                            lambda qnn, m, name, unit_id=0, **k: seen.append(("layer", name, unit_id, k["iters"])))
        monkeypatch.setattr(main2, "block_reconstruction",
                            lambda qnn, m, name, unit_id=0, **k: seen.append(("block", name, unit_id, k["iters"])))
        monkeypatch.setattr(ocalib, "reconstruct", lambda qnn, m, uid, name, cali, **k: oseen.append(
            ("block" if isinstance(m, owrap.BaseQuantBlock) else "layer", name, uid, k["iters"])))
        traces = main2.recon_model(p, iters=7)
        otraces = ocalib.recon_model(o, None, iters=7)
        assert seen == oseen and list(traces) == list(otraces) and len(seen) >= 20, arch
        assert [s[2] for s in seen] == sorted(s[2] for s in seen)                  # unit ids follow the walk
        if arch == "cheng2020-attn":
            assert any(k == "block" for k, *_ in seen) and any(k == "layer" for k, *_ in seen)
    # the image output layer the reference patches by hand (main2.py:256-263)
    assert main2.output_layer(p, True) is p.model.g_s[-1][0]


def test_bench_reference_arm_json_contract(monkeypatch, capsys):
    """`bench.py --impl reference` (the host-CPU arm): one JSON line with the contract's keys, the calibration baseline
    and both forward baselines (verbatim per-channel loop / vectorised).  Shrunk codec so it runs in seconds."""
    import argparse
    import json
    import bench
    monkeypatch.setattr(bench, "N_CH", 8)
    monkeypatch.setattr(bench, "M_CH", 12)
    monkeypatch.setattr(bench, "PATCH", 64)
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(argparse.Namespace(steps=1, warmup=0))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "fwd_mpx_s"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "calib imgs/s" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["units"] == 20
    assert cb["fwd_mpx_s_verbatim"] > 0 and cb["fwd_mpx_s_vectorised"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("RDO-PTQ AdaRound calibration sweep")
    monkeypatch.setenv("RANK", "1")                       # other ranks exit without work or output
    bench.run_reference(argparse.Namespace(steps=1, warmup=0))
    assert capsys.readouterr().out == ""


def test_conv_engine_plan_pairs_and_options():
    """b200lic_conv_plan_info (host-side planning, no device): the CTA-pair form is taken for every eligible shape (output-
    channel tile a multiple of 32, at least two pixel tiles, not GDN), its grid is an even number of CTAs of at most the
    SM count, one pixel tile per CTA with double-buffered accumulators; options "pair" / "streamk" switch the forms."""
    from rdo_ptq_b200 import _lib, ops
    L = _lib.lib()

    def plan(shape, w, st, tr=False, gdn=False):
        d = ops.gdn_desc(shape, False) if gdn else ops.conv_desc(shape, w, st, w[2] // 2, tr, st - 1 if tr else 0)
        return ops.conv_plan_info(d, tr)

    try:
        p = plan((8, 192, 128, 128), (192, 192, 5, 5), 2)                    # g_a.2 of the benchmark
        assert p["eligible"] and p["pair"] == 1 and p["BN"] == 192 and p["MT"] == 1 and p["stream_k"] == 0
        assert p["m_tiles"] == 256 and p["items"] == 128 and p["grid"] == 148 and p["tmem_cols"] == 512
        p2k = plan((1, 192, 768, 1024), (192, 192, 5, 5), 2)
        assert p2k["pair"] == 1 and p2k["items"] == 768 and p2k["grid"] == 148
        pt = plan((8, 192, 32, 32), (192, 192, 5, 5), 2, tr=True)           # four sub-pixel phases x 64 tiles
        assert pt["pair"] == 1 and pt["items"] == 4 * 32 and pt["grid"] % 2 == 0
        narrow = plan((1, 64, 24, 16), (96, 64, 3, 3), 1)                    # few pixels: the cost model spreads the
        assert narrow["pair"] == 0 and narrow["BN"] == 16 and narrow["n_tiles"] == 6     # K loop over 16-channel tiles
        os.environ["B200LIC_TC_BN"] = "256"                                  # (experiments knob: widest tile)
        odd = plan((1, 64, 24, 16), (96, 64, 3, 3), 1)                       # 3 pixel tiles: the last pair has one
        del os.environ["B200LIC_TC_BN"]
        assert odd["pair"] == 1 and odd["BN"] == 96 and odd["m_tiles"] == 3 and odd["items"] == 2 and odd["grid"] == 4
        small = plan((2, 16, 64, 64), (16, 16, 5, 5), 2)                     # 16-channel tile: single CTA
        assert small["pair"] == 0 and small["grid"] == small["items"]
        assert plan((8, 192, 128, 128), None, 1, gdn=True)["pair"] == 0      # GDN's K loop is six blocks
        assert L.b200lic_set_option(b"pair", 0) == 0
        single = plan((8, 192, 128, 128), (192, 192, 5, 5), 2)
        assert single["pair"] == 0 and single["grid"] <= 148 and single["m_tiles"] == 256
        assert L.b200lic_set_option(b"pair", 2) == 0 and L.b200lic_set_option(b"streamk", 2) == 0
        sk = plan((8, 192, 128, 128), (192, 192, 5, 5), 2)                   # stream-K over the pairs (tests only)
        assert sk["pair"] == 1 and sk["stream_k"] == 1 and sk["grid"] == 148
        assert L.b200lic_set_option(b"nonsense", 1) != 0
    finally:
        os.environ.pop("B200LIC_TC_BN", None)
        L.b200lic_set_option(b"pair", 1)
        L.b200lic_set_option(b"streamk", 1)


def test_quant_model_rewrites_swin_blocks_like_the_reference():
    """QuantModel's graph rewrite (quant_model.py:23-66) over a model that holds convolutions and an RSTB: the RSTB becomes
    a QuantRSTB (specials), its Linear / LayerNorm children QuantModules, with the same module tree as the reference's own
    QuantModel builds from its own RSTB (imported through the shim when /root/reference is present).  Construction only."""
    import torch.nn as nn
    from rdo_ptq_b200 import codec, quantization as Q

    def tiny(rstb_cls, conv=nn.Conv2d):
        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.g_a0 = conv(3, 32, 5, stride=2, padding=2)
                self.g_a1 = rstb_cls(dim=32, input_resolution=(8, 8), depth=2, num_heads=4, window_size=4, mlp_ratio=2.)
                self.g_a2 = conv(32, 32, 3, stride=2, padding=1)
        return Net()

    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    q = Q.QuantModel(tiny(codec.RSTB), wq, aq)
    assert isinstance(q.model.g_a1, Q.QuantRSTB) and isinstance(q.model.g_a0, Q.QuantModule)
    blk = q.model.g_a1.residual_group.blocks[1]
    assert isinstance(blk, Q.QuantSwinTransformerBlock) and blk.geom.shift_size == 2 and blk.geom.attn_mask is not None
    assert isinstance(blk.attn.qkv, Q.QuantModule) and blk.attn.qkv.is_linear and blk.norm1.if_layer_norm
    assert blk.mlp.fc1.disable_act_quant and not blk.mlp.fc2.disable_act_quant

    def tree(model, QM, QB):
        return [(n.replace(".geom", "@"), type(m).__name__) for n, m in model.named_modules()
                if isinstance(m, (QM, QB)) and ".geom." not in n]

    mine = tree(q, Q.QuantModule, Q.BaseQuantBlock)
    assert len([t for t in mine if t[1] == "QuantModule"]) == 2 + 12
    from oracle import _ref_shim as S
    if S.available():
        TO = S.import_task_oriented()
        from quantization import quant_layer as r_ql, quant_block as r_qb
        from models import layers as r_layers
        r = TO.QuantModel(tiny(r_layers.RSTB), wq, aq)
        assert tree(r, r_ql.QuantModule, r_qb.BaseQuantBlock) == mine


def test_lu2022_graph_and_its_rewrite_match_the_reference():
    """codec.NIC (the Lu2022 graph, nic_cvt.py:21-330): same state-dict keys and shapes as the reference's own NIC, and
    QuantModel rewrites it into the same tree of QuantModules / QuantRSTBs (imported through the shim).  Construction only:
    the composed forward has not been run on hardware (see the header of rdo_ptq_b200/codec/nic.py)."""
    import warnings
    from rdo_ptq_b200 import codec, quantization as Q
    from oracle import _ref_shim as S
    cfg = dict(height=128, width=128, in_chans=3, embed_dim=32, latent_dim=48, window_size=4, mlp_ratio=2., qkv_bias=True,
               qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, use_checkpoint=False)
    wq = dict(n_bits=8, channel_wise=True, scale_method="max")
    aq = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
    m = codec.NIC(cfg)
    assert m.g_a1.residual_group.blocks[1].shift_size == 2 and m.h_a1.residual_group.blocks[0].window_size == 2
    q = Q.QuantModel(codec.NIC(cfg), wq, aq)
    mine = [(n, type(mod).__name__) for n, mod in q.named_modules()
            if isinstance(mod, (Q.QuantModule, Q.BaseQuantBlock)) and ".geom." not in n]
    assert sum(1 for _, t in mine if t == "QuantRSTB") == 12
    if not S.available():
        return
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        TO = S.import_task_oriented()
        from quantization import quant_layer as r_ql, quant_block as r_qb
        from models import nic_cvt as r_nic
        r = r_nic.NIC(cfg)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in r.state_dict().items()}
        rq = TO.QuantModel(r_nic.NIC(cfg), wq, aq)
        ref = [(n, type(mod).__name__) for n, mod in rq.named_modules()
               if isinstance(mod, (r_ql.QuantModule, r_qb.BaseQuantBlock))]
        assert mine == ref
