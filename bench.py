#!/usr/bin/env python
"""bench.py -- headline benchmark of the RDO-PTQ hot path on B200 (contract: the task brief / DESIGN.md section 6).

Workload (BASELINE.json configs[1]): W8 RDO-PTQ (AdaRound) calibration of Minnen2018 mean-scale (mbt2018-mean,
N=192, M=320, random init) on synthetic 256x256 calibration patches.  One *step* = one AdaRound iteration
(batch pick + QDrop mix, soft-quantised weight, forward, rec+task loss, wgrad, [all-reduce], STE/regulariser/Adam)
on EVERY reconstruction unit of the model (20 QuantModules), per-GPU batch 8 (weak scaling: global batch = 8*N).
metric `calib imgs/s` = units * global_batch / step time = sum_layers(iters * global_batch) / sum_layers(loop time)
(SURVEY.md 8(d)).

`value` is the REFERENCE'S PROCEDURE: the units run one after the other (one stream; with N > 1 every unit's all-reduce +
Adam tail completes before the next unit starts, nothing of one unit is hidden under another -- SURVEY 8(e): "no
layer-parallelism").  The schedule that overlaps independent units on three streams is reported beside it as
`value_overlapped` (a throughput proxy: it treats the units as independent problems).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the oracle port of the reference on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH, N_CH, M_CH, GAIN = "mbt2018-mean", 192, 320, 1.2
PATCH, PER_GPU_BATCH, POOL = 256, 8, 64
STRONG_GLOBAL_BATCH = 64
WQ = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)
CALIB = dict(iters=20000, weight=0.01, b_range=(20, 2), warmup=0.2, input_prob=0.5)
SEQUENTIAL = dict(n_streams=1, overlap_update=False)


def workload_name(n_units):
    """The same string in both arms' `config.workload` (BASELINE.json configs[1])."""
    return (f"RDO-PTQ AdaRound calibration sweep, {ARCH} N={N_CH} M={M_CH} random-init, {n_units} units x batch "
            f"{PER_GPU_BATCH}/GPU of {PATCH}x{PATCH} patches, W8 per-channel, QDrop 0.5")


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written).  `tf` = the burst cuBLAS bf16 figure, the one for a
    kernel timed in isolation as the roofline kernels below are; `tf_sustained` = the seconds-long figure under the power
    cap, reported next to it.  Fallback (file absent): the profiling recipe's 6.65 TB/s / 1.59 PFLOP/s burst."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        tf = d.get("bf16_tflops", 1590.0)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=tf, tf_sustained=d.get("bf16_tflops_sustained", tf), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named op from the committed
    `ncu --set full` capture of this build (profiles/traffic.json, written by scripts/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def layer_macs(session_units, caches):
    """Algorithmic MACs per sample of each unit's forward: conv Ho*Wo*Cout*Cin*k*k, tconv
    Hin*Win*Cin*Cout*k*k, GDN H*W*C^2; a block = the sum over its QuantModules (hooked shapes)."""
    from rdo_ptq_b200.quantization.session import unit_macs
    return {name: unit_macs(u, caches[name]) for name, u in session_units}


def graph_time_ms(fn, flush, reps=10, R=5):
    """Device time of `fn`'s kernels with a cold L2 and no host gaps: graph A = R x (256 MB flush, fn), graph B = R x
    flush; median over `reps` replays of (A - B) / R.  CUDA events on the launching stream."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def timed(body):
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            for _ in range(R):
                body()
        out = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cg.replay()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        out.sort()
        return out[len(out) // 2]

    def both():
        flush.zero_()
        fn()
    return (timed(both) - timed(flush.zero_)) / R


# ------------------------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch.distributed as dist
    from rdo_ptq_b200 import codec, synth, ops, _lib, evaluate as E
    from rdo_ptq_b200.quantization import QuantModel
    from rdo_ptq_b200.quantization.session import CalibrationSession

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG / NCCL_DEBUG_FILE are left exactly as the caller set them (the driver counts ranks from NCCL's own
        # log); the JSON line is printed LAST, after the process group is gone, so it stays the final line of stdout
        dist.init_process_group("nccl", device_id=dev)
    _lib.call("actq_stats_init", ops._p(torch.empty(2, dtype=torch.int32, device=dev)), 1, stream=None)  # arch gate early
    pk = peaks()

    def build(arch=ARCH, gain=GAIN, wq=WQ, aq=AQ, **kw):
        torch.manual_seed(1005)
        kw = kw or (dict(N=N_CH) if arch == "cheng2020-attn" else dict(N=N_CH, M=M_CH))
        m = codec.ARCHS[arch](**kw).eval()
        synth.init_weights(m, gain=gain)
        m.to(dev)
        with torch.no_grad():
            m(synth.calibration_patches(1, 64).to(dev))         # one FP forward first: bakes the MaskedConv2d mask (Q5)
        return QuantModel(m, wq, aq, is_cheng=(arch == "cheng2020-attn")).eval()

    sampler = ClockSampler(local) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def timed(session, steps, warmup):
        for _ in range(max(warmup, session.graph_warmup + 1)):      # eager warm-up sweeps + the graph-capture sweep
            session.sweep()
        barrier()
        l0 = session.launch_total()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            session.sweep()
        e1.record()
        barrier()
        return maxr(e0.elapsed_time(e1)), session.launch_total() - l0

    def calib_leg(batch, pool, steps, warmup, session_kw, arch=ARCH, gain=GAIN, want_session=False):
        cali = synth.calibration_patches(pool, PATCH, seed=1005 + rank).to(dev)     # each rank: its own shard of the pool
        qnn = build(arch, gain)
        sess = CalibrationSession(qnn, cali, batch_size=batch, host_caches=False, **CALIB, **session_kw)
        ms, launches = timed(sess, steps, warmup)
        out = dict(value=len(sess.units) * batch * world * steps / (ms / 1e3), ms_per_step=ms / steps,
                   units=len(sess.units), per_gpu_batch=batch, launches=launches)
        if want_session:
            return out, sess
        del sess, qnn
        torch.cuda.empty_cache()
        return out

    skw = json.loads(os.environ.get("B200LIC_SESSION_KW", "{}"))       # experiments only
    if sampler:
        sampler.start()
    # ---- headline: sequential walk, caches resident in HBM ------------------------------------------------------------
    seq, sess = calib_leg(PER_GPU_BATCH, POOL, args.steps, args.warmup, dict(SEQUENTIAL, **skw), want_session=True)
    n_units = seq["units"]
    macs = layer_macs(sess.units, sess.caches)

    # ---- dominant kernel roofline: the forward GEMM of g_a.2 (192->192 5x5 s2 conv) as the calibration iteration issues it:
    # tc2_gather_gemm_kernel on PREPARED operands (activation operand staged by stage_mix, weight operand by quant_pack) --
    flush = torch.empty(64 * 1024 * 1024, device=dev)       # 256 MB > 126 MB L2
    top = "g_a.2"
    u = dict(sess.units)[top]
    q_in = sess.caches[top][0][:PER_GPU_BATCH].contiguous()
    wq_ = u.weight_quantizer
    d = ops.conv_desc(q_in.shape, u.weight.shape, u.fwd_kwargs["stride"], u.fwd_kwargs["padding"])
    ws = ops._workspace(d, ops.fwd_op(False), dev)
    slot = ops.conv_x_slot(d, False, ws)
    ops.stage_mix_sched(q_in, q_in, None, PER_GPU_BATCH, 1.0, 0, 1, 0, None, slot)
    packed = ops.quant_pack_weights(u.weight.data, wq_.alpha.data, wq_.delta, wq_.zero_point, wq_.axis, wq_.n_levels, True,
                                    d, False)
    y_buf = torch.empty((d.N, d.Cout, d.Ho, d.Wo), device=dev)
    k_ms = graph_time_ms(lambda: ops.conv_fwd_packed(None, packed, d, False, bias=u.bias.data, ws=ws, y=y_buf), flush)
    flops = 2.0 * macs[top] * PER_GPU_BATCH
    ach = flops / (k_ms * 1e-3) / 1e12
    # the un-fused op of round 1 (NHWC split + weight pack + GEMM), for continuity
    w_soft = ops.adaround_fwd(u.weight.data, wq_.alpha.data, wq_.delta, wq_.zero_point, wq_.axis, wq_.n_levels, True)
    op_ms = graph_time_ms(lambda: ops.conv2d_raw(q_in, w_soft, u.bias.data, d), flush)
    alg_bytes = 2.0 * 2 * q_in.numel() + 2.0 * 2 * u.weight.numel() + 4.0 * y_buf.numel()
    roof = {"bound": "tensor", "kernel": "tc2_gather_gemm_kernel<pair>: conv_fwd g_a.2 192->192 5x5 s2 @[8,192,128,128] on "
                                         "prepared operands (the forward GEMM of the fused AdaRound iteration; CTA-pair "
                                         "form, tcgen05 cta_group::2, M = 256 per MMA)",
            "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"],
            "peak_sustained": pk["tf_sustained"], "frac_of_sustained": ach / pk["tf_sustained"],
            "traffic": ncu_traffic("tc2_gather_gemm_kernel g_a.2"), "algorithmic_bytes": alg_bytes,
            "peak_source": pk["src"], "ms_per_launch": k_ms,
            "unfused_op_ms": op_ms, "unfused_op_tflops": flops / (op_ms * 1e-3) / 1e12,
            "note": "kernel timed in isolation: graph replay, L2 flushed, flush subtracted; `peak` = measured BURST bf16 "
                    "figure; 3 bf16 MMA passes per product (fp32-accurate split) bound frac at 1/3; algorithmic bytes = "
                    "split-bf16 x (4 B/elem) + packed weights (4 B/elem) + fp32 y; `traffic` = dram bytes of this kernel "
                    "from the committed ncu --set full capture (profiles/r2p_ncu_full_pair_raw.csv: tensor pipe 67 % of elapsed / "
                    "81 % of active cycles, cold cache); `unfused_op_*` = round 1's op (NHWC split + pack + GEMM)"}
    del sess
    torch.cuda.empty_cache()

    hbm_rows = []
    if not args.skip_extra:
        hbm_rows = hbm_rooflines(ops, dev, flush, pk)

    # ---- the overlapped schedule and the strong-scaling run ------------------------------------------------------------
    ovl = calib_leg(PER_GPU_BATCH, POOL, args.steps, args.warmup, dict(n_streams=3, **skw))
    strong = None
    if not args.skip_extra and STRONG_GLOBAL_BATCH % world == 0:
        b = STRONG_GLOBAL_BATCH // world
        strong = calib_leg(b, max(POOL, b), max(3, args.steps // 4), 3, dict(SEQUENTIAL))
        strong["global_batch"] = STRONG_GLOBAL_BATCH
        strong["scaling"] = "strong"

    # ---- end-to-end runs through the public API with HOST buffers (e2e) -------------------------------------------------
    # streaming mode: the calibration images live in pinned host memory; every step copies its batch of images
    # host->device, recomputes every unit's (quant_in, fp_in, fp_out) with two captured forwards, runs the sweep and reads
    # the per-unit losses back.
    def e2e_run(mode, session_kw):
        cali = synth.calibration_patches(POOL, PATCH, seed=1005 + rank)
        qnn2 = build()
        sess2 = CalibrationSession(qnn2, cali if mode == "stream" else cali.to(dev), batch_size=PER_GPU_BATCH,
                                   host_caches=mode, **CALIB, **session_kw)
        # loss read-back every step (rec/task/round per unit): a non-blocking copy into pinned memory whose values are
        # consumed one step later, so the host queues step k+1 while step k runs (B200LIC_E2E_LAG=0: blocking read)
        lag = os.environ.get("B200LIC_E2E_LAG", "1") != "0"
        for _ in range(max(3, args.warmup)):          # >= graph_warmup eager sweeps + the capture sweep
            sess2.sweep()
            sess2.losses(lag=lag)
        barrier()
        sess2.h2d_bytes = 0
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            sess2.sweep()
            sess2.losses(lag=lag)
            d2h += 4 * 3 * n_units
        barrier()
        e2e_s = maxr(time.perf_counter() - t0)
        out = {"value": n_units * PER_GPU_BATCH * world * args.steps / e2e_s, "unit": "imgs/s",
               "h2d_bytes_per_step": sess2.h2d_bytes // args.steps, "d2h_bytes_per_step": d2h // args.steps}
        del sess2
        torch.cuda.empty_cache()
        return out

    e2e = e2e_run("stream", dict(SEQUENTIAL))
    e2e["mode"] = ("streaming calibration, units one after the other: batch images H2D each step, unit inputs/targets "
                   "recomputed on device, per-unit losses D2H")
    e2e_ovl = e2e_run("stream", dict(n_streams=3))
    e2e_ovl["mode"] = "the same with the units overlapped on three streams"

    # ---- second half of the metric: W8A8 evaluation forward Mpx/s -------------------------------------------------------
    # Images are independent, so evaluation shards them over the ranks.  768x512 (Kodak shape, configs 1/4)
    # and 2K CLIC shape (1365x2048 padded to 1536x2048, config 5; Mpx/s counts the unpadded pixels): every rank times its
    # own image.  `eval_2k_41`: the 41-image sweep of config 5 through evaluate.evaluate -- images round-robin over the
    # ranks, PSNR / bpp per image, one 3-number all-reduce.
    fwd = {}
    if not args.skip_fwd:
        fwd = forward_legs(args, build, E, synth, dev, rank, world, barrier, maxr)

    # ---- config 3: Cheng2020-attention calibration (blocks reconstructed jointly, masked context model in parallel) ----
    cheng = None
    if not args.skip_extra:
        cheng = calib_leg(PER_GPU_BATCH, 16, max(3, args.steps // 4), 3, dict(SEQUENTIAL), arch="cheng2020-attn", gain=0.6)
        cheng["workload"] = "cheng2020-attn N=192 W8 AdaRound calibration sweep (config 3), sequential, batch 8/GPU"

    clocks = sampler.summary() if sampler else None
    cpu = parity = eager = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_baseline(steps=1, batch=PER_GPU_BATCH, units_subset=None)
        if not args.skip_fwd:
            cpu.update(cpu_fwd_baseline())
            parity = parity_leg(dev)
        if not args.skip_extra:
            eager = torch_eager_gpu(dev, args.steps)
    if rank == 0:
        line = {"metric": "calib imgs/s", "value": seq["value"], "unit": "imgs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": seq["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(n_units), "per_gpu_batch": PER_GPU_BATCH, "units": n_units,
                           "schedule": "sequential: one unit after the other on one stream (the reference's procedure); "
                                       "N>1: each unit's all-reduce + Adam completes before the next unit starts",
                           "pool": f"{POOL} patches per rank (the reference draws from 1024; the pool only sets which "
                                   f"cached rows a batch gathers -- per-iteration work is independent of it once the "
                                   f"caches exceed L2, and all 20 units' caches must be resident at once here: "
                                   f"{POOL} x 0.21 GB)",
                           "l2_policy": "inputs larger than L2 (unit caches total 13 GB; a different unit each call)",
                           "engine": os.environ.get("B200LIC_ENGINE", "auto"),
                           "launch": "one CUDA graph per unit per iteration (device-resident schedule)"},
                "value_overlapped": ovl["value"], "ms_per_step_overlapped": ovl["ms_per_step"],
                "strong_scaling_gb64": strong, "e2e": e2e, "e2e_overlapped": e2e_ovl, "gpu_launches": seq["launches"],
                "clocks": clocks, "roofline": roof, "roofline_hbm": hbm_rows, "cpu_baseline": cpu, "parity": parity,
                "torch_eager_gpu": eager, "config3_cheng2020_calib": cheng,
                "gflop_per_step": 2 * 2 * sum(macs.values()) * PER_GPU_BATCH / 1e9,
                "tflops_step": 2 * 2 * sum(macs.values()) * PER_GPU_BATCH / 1e12 / (seq["ms_per_step"] / 1e3),
                # layers that asked for the tensor-core engine and ran on the exact-fp32 SIMT engine instead (must be 0)
                "simt_fallbacks": _lib.simt_fallback_count()}
        line.update(fwd)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world > 1:
            time.sleep(1.0)                    # let the other ranks' teardown messages drain first
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


def hbm_rooflines(ops, dev, flush, pk):
    """HBM-bound kernels of the path at batches that leave the launch-latency regime, WITH the likelihood tensors
    materialised as the product forward does (K9 20 B/elem, K10 12 B/elem, K11 12 B/elem, K8 12 B/elem,
    GDN forward 8 B/elem / 12 B/elem with the norm output calibration keeps for backward)."""
    g = torch.Generator().manual_seed(1005)
    rows = []

    def add(name, fn, nbytes, batch):
        ms = graph_time_ms(fn, flush)
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": gbs / pk["hbm"], "ms_per_launch": ms, "algorithmic_bytes": nbytes, "batch": batch})

    NB = 24
    y = (torch.randn(NB, 320, 96, 128, generator=g) * 3).to(dev)
    par = torch.randn(NB, 640, 96, 128, generator=g).to(dev)
    sc, mu = par.chunk(2, 1)
    add("K9 gaussian_lik fwd + lik (20 B/elem)", lambda: ops.gaussian_lik(y, sc, mu, want_lik=True), 20.0 * y.numel(),
        f"[{NB},320,96,128]")
    del y, par, sc, mu
    z = (torch.randn(NB * 8, 192, 24, 32, generator=g) * 2).to(dev)
    pkd = torch.randn(192, 58, generator=g).to(dev) * 0.5
    med = torch.zeros(192, device=dev)
    tab = ops.factorized_table(pkd, med)
    add("K10 factorized_lik fwd + lik (12 B/elem)", lambda: ops.factorized_lik(z, pkd, med, want_lik=True, table=tab),
        12.0 * z.numel(), f"[{NB * 8},192,24,32]")
    del z
    a = torch.randn(8, 192, 128, 128, generator=g).to(dev)
    b2 = torch.randn(8, 192, 128, 128, generator=g).to(dev)
    add("K11 lp_loss value + gradient (12 B/elem)", lambda: ops.lp_loss_fwd_bwd(a, b2), 12.0 * a.numel(), "[8,192,128,128]")
    add("K8 dynamic A8 stats + apply (12 B/elem)", lambda: ops.act_quant(a), 12.0 * a.numel(), "[8,192,128,128]")
    gam = (torch.rand(192, 192, generator=g) * 0.01 + 0.1 * torch.eye(192)).to(dev)
    bet = torch.ones(192).to(dev)
    dg = ops.gdn_desc(a.shape, False)
    xa = a.abs() + 0.1
    pk_g = ops.pack_weights(gam.view(192, 192, 1, 1), dg, False)
    yg = torch.empty_like(xa)
    keys = ops.act_quant_stats(xa)
    add("K3 GDN forward, one kernel (8 B/elem: x in, y out)", lambda: ops.gdn_fwd_fused(xa, pk_g, bet, False, y=yg),
        8.0 * a.numel(), "[8,192,128,128]")
    add("K3+K8 GDN forward with the deferred A8 quantiser of its input applied on chip (8 B/elem)",
        lambda: ops.gdn_fwd_fused(xa, pk_g, bet, False, pending=(keys, 8), y=yg), 8.0 * a.numel(), "[8,192,128,128]")
    add("K3 GDN forward, round-1 form: staged x^2 operand + conv engine (8 B/elem algorithmic)",
        lambda: ops.conv2d_raw(xa, gam.view(192, 192, 1, 1), bet, dg, gdn_x=xa), 8.0 * a.numel(), "[8,192,128,128]")
    add("K3 GDN forward + norm for backward (12 B/elem)",
        lambda: ops.conv2d_raw(xa, gam.view(192, 192, 1, 1), bet, dg, gdn_x=xa, want_norm=True), 12.0 * a.numel(),
        "[8,192,128,128]")
    return rows


def forward_legs(args, build, E, synth, dev, rank, world, barrier, maxr):
    out = {}
    qnn3 = build()
    res, gf = {}, None

    def w8a8(q, xp):
        q.set_quant_state(True, False)
        q(xp)
        for m in q.modules():
            if hasattr(m, "trained"):
                m.trained = True
        q.set_quant_state(True, True)
        last = q.model.g_s[-1]
        (last[0] if isinstance(last, torch.nn.Sequential) else last).set_quant_state(True, False)

    def time_forward(gf, xp, h, w_, reps=5):
        for _ in range(3):                 # eager pass, capture pass, first replay
            gf(xp)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            gf(xp)
        b.record()
        barrier()
        return world * reps * h * w_ / 1e6 / (maxr(a.elapsed_time(b)) / 1e3)

    with torch.no_grad():
        for tag, (h, w_) in (("768x512", (512, 768)), ("2k", (1365, 2048))):
            xp = E.pad(synth.synthetic_image(h, w_, seed=1005 + rank).to(dev), 256)
            if gf is None:
                w8a8(qnn3, xp)
                gf = E.GraphedForward(qnn3)
            res[tag] = time_forward(gf, xp, h, w_)
        out["fwd_mpx_s"], out["fwd_mpx_s_2k"] = res["768x512"], res["2k"]
        if not args.skip_extra:
            # config 5: 41 synthetic 2K CLIC-shape images sharded round-robin, PSNR / bpp reduced with one all-reduce
            imgs = [None] * 41
            mine = list(range(rank, 41, world))
            base = [synth.synthetic_image(1365, 2048, index=i).to(dev) for i in mine[:2]]    # two distinct images, reused
            for j, i in enumerate(mine):
                imgs[i] = base[j % len(base)]
            for i in range(41):
                if imgs[i] is None:
                    imgs[i] = base[0]                          # other ranks' slots: never touched by this rank
            E.evaluate(qnn3, imgs[:world], shard=True)          # warm-up: capture the 2K graph
            barrier()
            t0 = time.perf_counter()
            r = E.evaluate(qnn3, imgs, shard=True)
            barrier()
            dt = maxr(time.perf_counter() - t0)
            out["eval_2k_41"] = {"mpx_s": 41 * 1365 * 2048 / 1e6 / dt, "seconds": dt, "images": r["count"],
                                 "psnr": r["psnr"], "bpp": r["bpp"],
                                 "note": "config 5: evaluate.evaluate over 41 2K images, round-robin over the ranks, "
                                         "per-image PSNR + bpp read back, one 3-number all-reduce; wall clock, max "
                                         "over ranks"}
    if not args.skip_extra and rank == 0:
        out["entropy_coding"] = coding_leg(qnn3, E, synth, dev)
    del qnn3, gf
    torch.cuda.empty_cache()
    if not args.skip_extra:
        from rdo_ptq_b200.quantization.quantizer import UniformAffineQuantizer as PUAQ
        with torch.no_grad():
            xp = E.pad(synth.synthetic_image(512, 768, seed=1005 + rank).to(dev), 256)
            q = build("cheng2020-attn", 0.6)
            w8a8(q, xp)
            out["fwd_mpx_s_cheng2020_w8a8"] = time_forward(E.GraphedForward(q), xp, 512, 768, reps=3)
            del q
            PUAQ.act_bits_follow_n_bits = True                 # config 4: W10A10 (additive switch)
            try:
                q = build("cheng2020-attn", 0.6, wq=dict(WQ, n_bits=10), aq=dict(AQ, n_bits=10))
                w8a8(q, xp)
                out["fwd_mpx_s_cheng2020_w10a10"] = time_forward(E.GraphedForward(q), xp, 512, 768, reps=3)
                del q
            finally:
                PUAQ.act_bits_follow_n_bits = False
        torch.cuda.empty_cache()
    return out


def coding_leg(qnn, E, synth, dev, reps=5):
    """Next row N2: real strings of the W8A8 codec's latents.  compress() = analysis transforms + symbol/index kernel +
    chunked rANS encode (sizes pass, prefix sum, write pass, D2H of the strings); decompress() = H2D + decode + synthesis.
    Wall clock (the strings cross the PCIe bus by definition).  `coder_msym_s`: the rANS kernels alone, strings on the device.
    `cpu_port_msym_s`: the oracle's sequential restatement (pure Python, 1 core) on a 20k-symbol sample of the same latents
    -- compressai's C++ coder is not installed here, so no faster CPU arm exists to time."""
    from rdo_ptq_b200.codec import coding
    from oracle import rans as R
    res = {}
    model = qnn.model
    model.update()
    with torch.no_grad():
        for tag, (h, w_) in (("768x512", (512, 768)), ("2k", (1365, 2048))):
            xp = E.pad(synth.synthetic_image(h, w_, seed=1005).to(dev), 256)
            out = model.compress(xp)
            rec = model.decompress(out["strings"], out["shape"])
            fwd = qnn(xp)
            exact = bool(torch.equal(rec["x_hat"], fwd["x_hat"].clamp(0, 1)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = model.compress(xp)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(reps):
                model.decompress(out["strings"], out["shape"])
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            est = sum(float((-torch.log2(l)).sum()) for l in fwd["likelihoods"].values()) / (h * w_)
            # the coder alone on the y latents
            y = model.g_a(xp)
            z_hat = model.entropy_bottleneck.decompress(out["strings"][1], out["shape"])
            scales, means = model._scales_means(model.h_s(z_hat))
            gc = model.gaussian_conditional
            idx = gc.build_indexes(scales)
            sym, _ = coding.symbols_and_indexes(y, means=means)
            t = gc._coding_tables(dev)
            n = sym.numel()
            chunk = coding.DEFAULT_CHUNK
            n_chunks = (n + chunk - 1) // chunk
            sizes = torch.empty(n_chunks, dtype=torch.int32, device=dev)
            offs = torch.zeros(n_chunks + 1, dtype=torch.int32, device=dev)
            sy, ix = sym.reshape(-1), idx.reshape(-1)
            args_t = (E.ops._p(t.cdf), E.ops._p(t.cdf_length), E.ops._p(t.offset), t.stride)
            E.ops.call("rans_encode_sizes", E.ops._p(sy), E.ops._p(ix), n, chunk, *args_t, E.ops._p(sizes))
            torch.cumsum(sizes, 0, out=offs[1:])
            words = torch.empty(int(offs[-1].item()), dtype=torch.int32, device=dev)
            dec = torch.empty(n, dtype=torch.int32, device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(reps):
                E.ops.call("rans_encode_sizes", E.ops._p(sy), E.ops._p(ix), n, chunk, *args_t, E.ops._p(sizes))
                E.ops.call("rans_encode_write", E.ops._p(sy), E.ops._p(ix), n, chunk, *args_t, E.ops._p(offs), E.ops._p(words))
            ev[1].record()
            for _ in range(reps):
                E.ops.call("rans_decode", E.ops._p(words), E.ops._p(offs), n, chunk, E.ops._p(ix), *args_t, E.ops._p(dec))
            ev[2].record()
            torch.cuda.synchronize()
            res[tag] = {"round_trip_exact": exact and bool(torch.equal(dec, sy)),
                        "file_bpp": 8 * coding.string_bytes(out["strings"]) / (h * w_), "estimated_bpp": est,
                        "compress_ms": (t1 - t0) / reps * 1e3, "decompress_ms": (t2 - t1) / reps * 1e3,
                        "compress_mpx_s": h * w_ / 1e6 / ((t1 - t0) / reps), "decompress_mpx_s": h * w_ / 1e6 / ((t2 - t1) / reps),
                        "y_symbols": n, "chunks": n_chunks,
                        "coder_encode_msym_s": n / 1e6 / (ev[0].elapsed_time(ev[1]) / reps / 1e3),
                        "coder_decode_msym_s": n / 1e6 / (ev[1].elapsed_time(ev[2]) / reps / 1e3)}
            if tag == "768x512":
                cdf, cdf_len, off = t.host
                s_np, i_np = sy[:20000].cpu().numpy(), ix[:20000].cpu().numpy()
                t0 = time.perf_counter()
                wds = R.rans64_encode(s_np, i_np, cdf, cdf_len, off)
                t1 = time.perf_counter()
                R.rans64_decode(wds, len(s_np), i_np, cdf, cdf_len, off)
                t2 = time.perf_counter()
                res["cpu_port_msym_s"] = {"encode": len(s_np) / 1e6 / (t1 - t0), "decode": len(s_np) / 1e6 / (t2 - t1),
                                          "cores": 1, "kind": "port",
                                          "sample": "first 20000 y symbols of the 768x512 image, pure-Python restatement"}
    res["note"] = ("mbt2018-mean N=192 M=320 W8A8; strings = header + chunk table + one rans64 stream per 2048 symbols; "
                   "file_bpp counts every byte of every string")
    return res


# ------------------------------------------------------------------------------------------------------------- CPU arms
def _oracle_sweep(dev, batch, tf32=None):
    """The oracle port of the calibration sweep on `dev` (CPU, or the GPU through PyTorch's own kernels as the same-box
    eager comparator).  Returns (sweep(), n_units)."""
    from oracle import codec as ocodec, quant_wrap as owrap, quantizers as oq
    from rdo_ptq_b200 import synth
    torch.manual_seed(1005)
    m = ocodec.ARCHS[ARCH](N=N_CH, M=M_CH).eval()
    synth.init_weights(m, gain=GAIN)
    m.to(dev)
    qnn = owrap.QuantModel(m, WQ, AQ).eval()
    cali = synth.calibration_patches(batch, PATCH).to(dev)
    units = [(n, u) for n, u in qnn.model.named_modules() if isinstance(u, owrap.QuantModule) and u.org_weight is not None]
    store = {}
    hooks = [u.register_forward_hook(lambda _m, i, o, n=n: store.__setitem__(n, (i[0].detach(), o.detach())))
             for n, u in units]
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        qnn(cali)
    fp = dict(store)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali)
    qin = {n: v[0] for n, v in store.items()}
    for h in hooks:
        h.remove()
    opts = {}
    for n, u in units:
        u.weight_quantizer = oq.AdaRoundQuantizer(u.weight_quantizer, u.org_weight.data)
        u.weight_quantizer.soft_targets = True
        u.set_quant_state(True, False)
        opts[n] = torch.optim.Adam([u.weight_quantizer.alpha])
    g = torch.Generator(device=dev).manual_seed(1)

    def sweep():
        for n, u in units:
            keep = torch.rand(qin[n].shape, generator=g, device=dev) < CALIB["input_prob"]
            cur = torch.where(keep, qin[n], fp[n][0])
            opts[n].zero_grad()
            out = u(cur)
            loss = oq.lp_loss(out, fp[n][1]) + oq.lp_loss(out, fp[n][1])
            loss.backward()
            opts[n].step()

    return sweep, len(units)


def cpu_baseline(steps=1, batch=PER_GPU_BATCH, units_subset=None, warmup=0):
    """The oracle port of the reference loop on the host cores (all threads), same model / patches / batch."""
    torch.set_num_threads(os.cpu_count())
    sweep, n_units = _oracle_sweep(torch.device("cpu"), batch)
    for _ in range(warmup):
        sweep()
    t0 = time.perf_counter()
    for _ in range(steps):
        sweep()
    dt = time.perf_counter() - t0
    return {"value": n_units * batch * steps / dt, "unit": "imgs/s", "cores": os.cpu_count(), "kind": "port",
            "units": n_units,
            "sample": f"{steps} sweep(s) of the same {n_units}-unit AdaRound iteration at batch {batch} "
                      f"({PATCH}x{PATCH}), PyTorch-CPU fp32 oracle, {os.cpu_count()} threads",
            "seconds": dt}


def torch_eager_gpu(dev, steps):
    """Same-box GPU comparator (the bar is PyTorch eager running the same fake-quant graph): the oracle
    port moved to the GPU -- cuDNN / ATen kernels, autograd, torch.optim.Adam -- for the same sweep and the same W8A8
    forward, with TF32 off (fp32-accurate like this repo's engine) and on."""
    from oracle import codec as ocodec, quant_wrap as owrap
    from rdo_ptq_b200 import synth
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            sweep, n_units = _oracle_sweep(dev, PER_GPU_BATCH)
            for _ in range(3):
                sweep()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = max(3, steps // 2)
            a.record()
            for _ in range(k):
                sweep()
            b.record()
            torch.cuda.synchronize()
            out[f"calib_imgs_s_{tag}"] = n_units * PER_GPU_BATCH * k / (a.elapsed_time(b) / 1e3)
            del sweep
            torch.cuda.empty_cache()
            torch.manual_seed(1005)
            m = ocodec.ARCHS[ARCH](N=N_CH, M=M_CH).eval()
            synth.init_weights(m, gain=GAIN)
            m.to(dev)
            qnn = owrap.QuantModel(m, WQ, AQ).eval()
            for hw_tag, (h, w_) in (("768x512", (512, 768)), ("2k", (1536, 2048))):
                x = synth.synthetic_image(h, w_).to(dev)
                with torch.no_grad():
                    qnn.set_quant_state(True, False)
                    qnn(x)
                    for mod in qnn.modules():
                        if hasattr(mod, "trained"):
                            mod.trained = True
                    qnn.set_quant_state(True, True)
                    qnn.model.g_s[-1].set_quant_state(True, False)
                    qnn(x)
                    torch.cuda.synchronize()
                    a.record()
                    for _ in range(3):
                        qnn(x)
                    b.record()
                    torch.cuda.synchronize()
                px = (1365 * 2048) if hw_tag == "2k" else h * w_
                out[f"fwd_mpx_s_{hw_tag}_{tag}"] = 3 * px / 1e6 / (a.elapsed_time(b) / 1e3)
            del qnn, m
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["note"] = ("oracle port (.cuda()): PyTorch eager, cuDNN convolutions, vectorised activation quantiser; the "
                   "comparator SURVEY section 2 names.  Not a parity-equivalent at tf32 (misses the 1e-4 per-layer bar)")
    return out


def cpu_fwd_baseline(h=512, w=768, reps=1):
    """Second half of the metric on the host cores: the oracle's W8A8 evaluation forward (dynamic A8 on) of one
    768x512 image, once with the activation quantiser as the reference ships it (Python loop over the channels,
    quantizer.py:99-117: "verbatim") and once with the vectorised equivalent (same arithmetic; SURVEY 8(d) asks for
    both so the speed-up is not inflated by interpreter overhead)."""
    from oracle import codec as ocodec, quant_wrap as owrap, quantizers as oq
    from rdo_ptq_b200 import synth
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1005)
    m = ocodec.ARCHS[ARCH](N=N_CH, M=M_CH).eval()
    synth.init_weights(m, gain=GAIN)
    qnn = owrap.QuantModel(m, WQ, AQ).eval()
    x = synth.synthetic_image(h, w)
    out = {}
    with torch.no_grad():
        qnn.set_quant_state(True, False)
        qnn(x)                                              # weight ranges
        for mod in qnn.modules():
            if hasattr(mod, "trained"):
                mod.trained = True
        qnn.set_quant_state(True, True)
        qnn.model.g_s[-1].set_quant_state(True, False)
        for tag, loop in (("vectorised", False), ("verbatim", True)):
            oq.UniformAffineQuantizer.act_verbatim_loop = loop
            try:
                t0 = time.perf_counter()
                for _ in range(reps):
                    qnn(x)
                out[tag] = reps * h * w / 1e6 / (time.perf_counter() - t0)
            finally:
                oq.UniformAffineQuantizer.act_verbatim_loop = False
    return {"fwd_mpx_s_verbatim": out["verbatim"], "fwd_mpx_s_vectorised": out["vectorised"],
            "fwd_sample": f"{reps} W8A8 forward(s) of one {w}x{h} image, {ARCH} N={N_CH}, oracle on {os.cpu_count()} threads"}


def parity_leg(dev):
    """Delta bpp / delta PSNR / worst per-layer relative error of the CUDA path against the oracle on the benchmark's
    own model at 768x512 (the oracle is the checker here, oracle/parity.py), on the tcgen05 engine and on the exact-fp32
    SIMT engine."""
    from oracle import parity as P
    rows = {}
    for eng in ("auto", "simt"):
        r = P.compare_forward(ARCH, dict(N=N_CH, M=M_CH), GAIN, (512, 768), dev, engine=eng, layer_checks=(eng == "auto"))
        rows[eng] = {k: r[k] for k in ("codes_equal", "worst_layer_rel_err", "worst_layer", "d_bpp_w", "d_psnr_w",
                                       "d_bpp_wa", "d_psnr_wa", "a8_flip_rate", "bpp_ref_wa", "psnr_ref_wa")}
    rows["bars"] = "codes bit-exact; per-layer 1e-4 relative; end-to-end 1e-3 bpp / 0.01 dB (north_star)"
    rows["case"] = f"{ARCH} N={N_CH} M={M_CH}, 768x512 synthetic image, W8 and W8A8, vs the pinned oracle"
    return rows


def run_reference(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    steps = max(1, min(args.steps, 20))          # one sweep is 1.5-4.5 s on 8-16 host cores: K = 20 stays within minutes
    warm = max(0, min(args.warmup, 5))
    r = cpu_baseline(steps=steps, warmup=warm)
    r.update(cpu_fwd_baseline())
    line = {"impl": "reference", "metric": "calib imgs/s", "value": r["value"], "unit": "imgs/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": steps, "warmup": warm,
            "ms_per_step": r["seconds"] / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(r["units"]), "per_gpu_batch": PER_GPU_BATCH, "units": r["units"],
                       "schedule": "sequential: one unit after the other (the reference's procedure)",
                       "arm": "oracle port of the reference loop (PyTorch-CPU fp32) on the host cores; a step is one "
                              "sweep of the same units at the same batch"},
            "cpu_baseline": r, "e2e": {"value": r["value"], "unit": "imgs/s", "h2d_bytes_per_step": 0,
                                       "d2h_bytes_per_step": 0},
            "fwd_mpx_s": r["fwd_mpx_s_verbatim"]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-fwd", action="store_true")
    ap.add_argument("--skip-extra", action="store_true",
                    help="headline legs only: no strong-scaling / Cheng2020 / 41-image / HBM-roofline / eager-GPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the "
                         "host-CPU arm)")
    run_cuda(args)


if __name__ == "__main__":
    main()
