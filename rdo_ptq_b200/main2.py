"""B200 drop-in for the calibration entry point task-oriented-PTQ/main2.py (parse_args :20-74, optimize_model :145-290,
recon_model :227-250): build the QuantModel, initialise the weight ranges with one forward, walk the reconstruction
units depth first (every unit sees the hardened units before it), then switch to W-n A-n for the final test.

Not carried (SURVEY section 8 marks them out of scope): checkpoint / dataset loaders, yaml config, TensorBoard logger,
MS-SSIM.  There are no pretrained checkpoints and no datasets on the box, so the codec is random-init under
`--seed` (the reference default 1005) and the calibration / test images are synthetic, or passed in as tensors by a
caller that has real ones (`optimize_model(args, model=..., cali_data=..., test_images=...)`).
"""
import argparse
import logging
import sys
import time

import torch
import torch.nn as nn

from . import codec, synth, evaluate as E
from .quantization import BaseQuantBlock, QuantModel, QuantModule, block_reconstruction, layer_reconstruction

ARCH_ALIASES = {"Minnen2018": "mbt2018-mean", "Cheng2020": "cheng2020-attn", "Balle2018": "bmshj2018-hyperprior"}


def parse_args(argv):
    """The reference's flags (main2.py:20-60) with its defaults; `--arch` additionally takes the codecs this package
    carries, and the synthetic-data knobs replace --config / --resume."""
    p = argparse.ArgumentParser(description='running parameters', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--seed', default=1005, type=int)
    p.add_argument('--quality', default=6, type=int, help='picks (N, M): 192/320 above the codec\'s split, else 128/192')
    p.add_argument('--batch_size', default=4, type=int)
    p.add_argument('--arch', default='Minnen2018', type=str,
                   choices=sorted(ARCH_ALIASES) + sorted(ARCH_ALIASES.values()))
    p.add_argument('--type', default='mse', type=str, choices=['mse'])
    p.add_argument('--lmbda', default=0.0483, type=float)
    p.add_argument('--save', type=str, default=None, help='path for torch.save(qnn) (main2.py:285-290)')
    p.add_argument('--n_bits_w', default=8, type=int)
    p.add_argument('--channel_wise', action='store_true')
    p.add_argument('--n_bits_a', default=8, type=int)
    p.add_argument('--act_quant', action='store_true')
    p.add_argument('--disable_8bit_head_stem', action='store_true')
    p.add_argument('--test_before_calibration', action='store_true')
    p.add_argument('--input_prob', default=0.5, type=float)
    p.add_argument('--lr', default=4e-5, type=float, help='accepted and ignored like the reference (layer_opt.py:254)')
    p.add_argument('--task_loss', default=2, help="p of the task lp_loss, or 'rd' for R + lambda*D")
    p.add_argument('--num_samples', default=12, type=int)
    p.add_argument('--iters_w', default=20000, type=int)
    p.add_argument('--weight', default=0.01, type=float)
    p.add_argument('--b_start', default=20, type=int)
    p.add_argument('--b_end', default=2, type=int)
    p.add_argument('--warmup', default=0.2, type=float)
    p.add_argument('--init', default='max', type=str, choices=['max', 'mse', 'gaussian', 'l1', 'l2'])
    # synthetic stand-ins for the reference's checkpoint and datasets
    p.add_argument('--N', default=None, type=int)
    p.add_argument('--M', default=None, type=int)
    p.add_argument('--patch', default=256, type=int, help='calibration patch size (config.yaml patch_size)')
    p.add_argument('--test_hw', default='512x768', type=str)
    p.add_argument('--n_test', default=2, type=int)
    p.add_argument('--gain', default=1.2, type=float, help='synthetic weight gain (keeps activations O(1))')
    args = p.parse_args(argv)
    if args.task_loss != 'rd':
        args.task_loss = float(args.task_loss)
    return args


def build_model(args, device):
    arch = ARCH_ALIASES.get(args.arch, args.arch)
    big = args.quality >= (4 if arch == "cheng2020-attn" else 6 if arch == "bmshj2018-hyperprior" else 5)
    N = args.N or (192 if big else 128)
    kw = dict(N=N) if arch == "cheng2020-attn" else dict(N=N, M=args.M or (320 if big else 192))
    torch.manual_seed(args.seed)
    model = codec.ARCHS[arch](**kw).eval()
    synth.init_weights(model, gain=args.gain)
    return model.to(device), arch


def output_layer(qnn, is_cheng):
    """The image output layer the reference patches by hand (main2.py:256-263, 273-278)."""
    return qnn.model.g_s[-1][0] if is_cheng else qnn.model.g_s[-1]


def recon_model(qnn: QuantModel, module: nn.Module = None, _counter=None, _prefix="", **kwargs):
    """main2.py:227-250: depth-first walk; QuantModules -> layer_reconstruction, BaseQuantBlocks -> block_reconstruction.
    Returns {unit path: [loss records]}.  `unit_id` (additive) numbers the units in walk order for the draw plan."""
    module = qnn if module is None else module
    counter = [0] if _counter is None else _counter
    traces = {}
    for name, m in module.named_children():
        full = f"{_prefix}.{name}" if _prefix else name
        if isinstance(m, (QuantModule, BaseQuantBlock)):
            kind = 'layer' if isinstance(m, QuantModule) else 'block'
            if m.ignore_reconstruction is True:
                logging.info('Ignore reconstruction of {} {}'.format(kind, name))
            else:
                logging.info('Reconstruction for {} {}'.format(kind, name))
                fn = layer_reconstruction if isinstance(m, QuantModule) else block_reconstruction
                # the unit's path inside the codec ("g_a.2"): what the R + lambda*D / coder task criteria continue from
                path = full[len("model."):] if full.startswith("model.") else full
                traces[full] = fn(qnn, m, name, unit_id=counter[0], unit_path=path, **kwargs)
            counter[0] += 1
        else:
            traces.update(recon_model(qnn, m, counter, full, **kwargs))
    return traces


def optimize_model(args, model=None, cali_data=None, test_images=None, device="cuda", plan=None, graph=True):
    """main2.py:145-290.  Returns (qnn, report); report holds the test results of every stage the reference logs."""
    report = {}
    if model is None:
        model, arch = build_model(args, device)
    else:
        arch = ARCH_ALIASES.get(args.arch, args.arch)
    is_cheng = arch == "cheng2020-attn"
    model.to(device).eval()
    if cali_data is None:
        cali_data = synth.calibration_patches(args.num_samples, args.patch)
    cali_data = cali_data.to(device)
    if test_images is None:
        h, w = (int(v) for v in args.test_hw.split("x"))
        test_images = synth.synthetic_images(args.n_test, h, w)
    test_images = [t.to(device) for t in test_images]

    def test(tag, net):
        report[tag] = E.evaluate(net, test_images, shard=False)
        logging.info('{}: psnr {:.4f} dB  bpp {:.5f}'.format(tag, report[tag]["psnr"], report[tag]["bpp"]))

    with torch.no_grad():
        model(cali_data[:1])        # one FP forward bakes the MaskedConv2d mask into weight.data (main2.py:169-171, Q5)
    if args.test_before_calibration:
        test('fp32', model)
    wq_params = {'n_bits': args.n_bits_w, 'channel_wise': args.channel_wise, 'scale_method': args.init}
    aq_params = {'n_bits': args.n_bits_a, 'channel_wise': args.channel_wise, 'scale_method': args.init,
                 'leaf_param': args.act_quant}
    qnn = QuantModel(model=model, weight_quant_params=wq_params, act_quant_params=aq_params, is_cheng=is_cheng)
    qnn.to(device).eval()
    if not args.disable_8bit_head_stem:
        qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    qnn.set_quant_state(True, False)
    t0 = time.time()
    with torch.no_grad():
        qnn(cali_data[:args.batch_size])                        # initialises every weight range (main2.py:194-198)
    report['init_time'] = time.time() - t0
    if args.test_before_calibration:
        test('w_nearest', qnn)
    kwargs = dict(cali_data=cali_data, batch_size=args.batch_size, iters=args.iters_w, weight=args.weight,
                  input_prob=args.input_prob, lr=args.lr, asym=True, b_range=(args.b_start, args.b_end),
                  warmup=args.warmup, act_quant=args.act_quant, opt_mode='mse', config=None, args=args, graph=graph,
                  lmbda=args.lmbda)
    if plan is not None:
        kwargs['plan'] = plan
    qnn.set_quant_state(weight_quant=True, act_quant=args.act_quant)
    output_layer(qnn, is_cheng).set_quant_state(True, False)
    t0 = time.time()
    report['losses'] = recon_model(qnn, **kwargs)
    torch.cuda.synchronize()
    report['calib_time'] = time.time() - t0
    qnn.set_quant_state(weight_quant=True, act_quant=False)
    test('w_opt', qnn.eval())
    qnn.set_quant_state(weight_quant=True, act_quant=True)
    output_layer(qnn, is_cheng).set_quant_state(True, False)
    test('wa_opt', qnn.eval())
    if args.save:
        torch.save(qnn, args.save)
    return qnn, report


def main(argv):
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    logging.info('task loss: {}  param init: {}  channel wise: {}  seed: {}  iterations: {}  batch_size: {}'.format(
        args.task_loss, args.init, args.channel_wise, args.seed, args.iters_w, args.batch_size))
    _, report = optimize_model(args)
    n_units = len(report['losses'])
    imgs = n_units * args.iters_w * args.batch_size
    logging.info('calibrated {} units in {:.1f} s ({:.0f} calib imgs/s)'.format(n_units, report['calib_time'],
                                                                                imgs / max(report['calib_time'], 1e-9)))


if __name__ == '__main__':
    main(sys.argv[1:])
