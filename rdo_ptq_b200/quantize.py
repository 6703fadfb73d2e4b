"""B200 drop-in for the one-shot PTQ entry point light-uniform-PTQ/quantize.py (parse_args :27-48, generator :85-114,
quantize_int8 :117-158): wrap the codec with the LU rules (uint8 per-channel weights, static Q8.8 activations, GDN left
in fp32), run one forward to materialise the integer weights, report PSNR / bpp, save the state dict.

The reference quantises TinyLIC from a downloaded checkpoint; neither is on the hot path.  BASELINE
config 1 runs the same rules on the Balle2018 scale hyperprior with random-init weights and synthetic images, which is
what this entry point builds unless a caller passes its own `model` / `images`.  `--type FP16` (a plain `.half()` cast,
quantize.py:161-186) is not carried.
"""
import argparse
import logging
import sys
import time

import torch

from . import codec, synth, evaluate as E
from .quant_int import QuantModel


def parse_args(argv):
    p = argparse.ArgumentParser(description='running parameters', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--seed', default=1005, type=int)
    p.add_argument('--save', type=str, default=None, help='path for torch.save(qnn.state_dict()) (quantize.py:155-157)')
    p.add_argument('--type', default='INT8', choices=['INT8'])
    p.add_argument('--n_bits_w', default=8, type=int)
    p.add_argument('--channel_wise', action='store_true', help='ignored like the reference: True is hard-coded (:144)')
    p.add_argument('--n_bits_a', default=16, type=int, help='ignored like the reference: activations are static Q8.8')
    p.add_argument('--act_quant', default=True)
    p.add_argument('--test_before_calibration', default=True, type=bool)
    p.add_argument('--init', default='max', choices=['max', 'mse'], help="weight range: 'max' (reference :144) or 'mse'")
    p.add_argument('--arch', default='bmshj2018-hyperprior', choices=sorted(codec.ARCHS))
    p.add_argument('--N', default=128, type=int)
    p.add_argument('--M', default=192, type=int)
    p.add_argument('--hw', default='512x768', type=str)
    p.add_argument('--n_test', default=2, type=int)
    p.add_argument('--gain', default=1.2, type=float)
    return p.parse_args(argv)


def generator(qnn, args, x):
    """quantize.py:85-114: one forward of a padded image with weight (and activation) quantisation on turns every
    wrapped layer's weight into uint8 codes + per-channel (delta, zero_point)."""
    qnn.set_quant_state(True, bool(args.act_quant))
    t0 = time.time()
    with torch.no_grad():
        qnn(E.pad(x, 64))
    torch.cuda.synchronize()
    logging.info('generate quantized model time: {}'.format(time.time() - t0))
    return qnn


def quantize_int8(args, model=None, images=None, device="cuda"):
    """quantize.py:117-158.  Returns (qnn, report)."""
    if model is None:
        torch.manual_seed(args.seed)
        kw = dict(N=args.N) if args.arch == "cheng2020-attn" else dict(N=args.N, M=args.M)
        model = synth.init_weights(codec.ARCHS[args.arch](**kw).eval(), gain=args.gain)
    model = model.to(device).eval()
    if images is None:
        h, w = (int(v) for v in args.hw.split("x"))
        images = synth.synthetic_images(args.n_test, h, w)
    images = [t.to(device) for t in images]
    report = {}
    # PSNR, MS-SSIM, bpp like validate_model (quantize.py:58-92); MS-SSIM needs sides > 160 px (five levels)
    ms = all(min(t.shape[-2:]) > 160 for t in images)
    keys = ('psnr', 'ms_ssim', 'bpp') if ms else ('psnr', 'bpp')
    fmt = '{}: psnr= {:.2f}; ms-ssim={:.4f}; bpp= {:.3f}' if ms else '{}: psnr= {:.2f}; bpp= {:.3f}'
    if args.test_before_calibration:
        report['fp32'] = E.evaluate(model, images, p=64, shard=False, ms_ssim=ms)
        logging.info(fmt.format('Full-precision model', *[report['fp32'][k] for k in keys]))
    wq_params = {'n_bits': args.n_bits_w, 'channel_wise': True, 'symmetric': False, 'scale_method': args.init}
    aq_params = {'channel_wise': False, 'symmetric': False, 'scale_method': 'max', 'leaf_param': True}
    qnn = QuantModel(model=model, weight_quant_params=wq_params, act_quant_params=aq_params)
    qnn.to(device).eval()
    qnn = generator(qnn, args, images[0])
    qnn.set_quant_state(weight_quant=True, act_quant=True)
    report['int8'] = E.evaluate(qnn, images, p=64, shard=False, ms_ssim=ms)
    logging.info(fmt.format('INT8', *[report['int8'][k] for k in keys]))
    if args.save:
        torch.save(qnn.state_dict(), args.save)
    return qnn, report


def main(argv):
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    quantize_int8(args)


if __name__ == '__main__':
    main(sys.argv[1:])
