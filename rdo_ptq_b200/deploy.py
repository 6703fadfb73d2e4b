"""True-integer deployment of a calibrated codec.

What the reference leaves open: light-uniform-PTQ saves `qnn.state_dict()` as INT8.pth (quantize.py:155-157) with the
uint8 codes in `weight` -- but the per-channel (delta, zero_point) are plain attributes of the quantiser objects, not
part of the state dict, and the float model refuses the uint8 tensors: "cannot reload INT8" (LU README.md:93-95);
task-oriented-PTQ pickles the whole QuantModel with fp32 weights next to alpha (main2.py:285-290).

`export_int8` writes what a decoder needs and nothing else: for every wrapped layer the hardened integer codes (uint8, or
int16 above 8 bit), per-channel delta / zero_point, the fp32 bias, and the parameters of the modules that stay in
floating point (entropy models; GDN under the LU rules).  `load_int8` rebuilds the codec and a QuantModel around it whose
weights are (code - zero_point) * delta: re-quantising those values returns the stored codes exactly ((n * delta) / delta
rounds back to the integer n for |n| < 2^16), so the loaded model's quantised forward is the calibrated model's, bit
for bit, and runs on the integer form of the conv engine (b200lic_quant_pack_weights integer mode: bf16-exact integer
weights, two tensor-core passes, delta applied per output channel in the epilogue).

Activations stay dynamic (task-oriented-PTQ, quantizer.py:81-121) or Q8.8 (light-uniform-PTQ): the per-channel activation
scale varies along the contraction axis, so it cannot be factored out of an integer GEMM; the integer operand is the weight.
"""
import io

import torch
import torch.nn as nn

from . import codec, ops
from .quantization import QuantModel, QuantModule
from .quantization.quantizer import AdaRoundQuantizer, UniformAffineQuantizer
from .quant_int import QuantModel as LUQuantModel, QuantModule as LUQuantModule

FORMAT = "b200lic-int8-v1"


def _codes(m):
    """(integer codes, delta [C], zero_point [C], axis, n_bits) of a wrapped layer's hardened weight."""
    if isinstance(m, LUQuantModule):
        q = m.weight_quantizer
        if m.weight.dtype != torch.uint8:
            raise RuntimeError("light-uniform model: run one forward first (quantize.py:85-114 materialises the codes)")
        return m.weight.data, q.delta.reshape(-1), q.zero_point.reshape(-1), q.channel_axis(m.weight), q.n_bits
    q = m.weight_quantizer
    if isinstance(q, AdaRoundQuantizer):
        if q.soft_targets:
            raise RuntimeError("AdaRound quantiser still has soft targets: finish the reconstruction before exporting")
        codes = q.codes(m.weight)
        axis = q.axis
    elif isinstance(q, UniformAffineQuantizer) and q.inited:
        codes = q.codes(m.weight)
        axis = q.channel_axis(m.weight)
    else:
        raise RuntimeError("weight quantiser is not initialised: run the range-initialising forward first")
    dt = torch.uint8 if q.n_levels <= 256 else torch.int16
    return codes.to(dt), q.delta.reshape(-1), q.zero_point.reshape(-1), axis, q.n_bits


def export_int8(qnn, arch: str, arch_kwargs: dict, path=None):
    """Integer deployment blob of a calibrated QuantModel (TO rules) or light-uniform QuantModel (LU rules)."""
    lu = isinstance(qnn, LUQuantModel)
    kinds = (LUQuantModule,) if lu else (QuantModule,)
    layers, wrapped_params = {}, set()
    for name, m in qnn.model.named_modules():
        if not isinstance(m, kinds) or getattr(m, "weight", None) is None:
            continue
        codes, delta, zp, axis, bits = _codes(m)
        layers[name] = dict(codes=codes.cpu(), delta=delta.detach().cpu().clone(), zero_point=zp.detach().cpu().clone(),
                            axis=axis, n_bits=bits, bias=None if m.bias is None else m.bias.detach().cpu().clone())
        for pn, _ in m.named_parameters():
            wrapped_params.add(f"{name}.{pn}")
    rest = {k: v.detach().cpu().clone() for k, v in qnn.model.state_dict().items()
            if k not in wrapped_params and not any(k.startswith(n + ".") for n in layers)}
    first = next(m for m in qnn.model.modules() if isinstance(m, kinds))
    # which wrapped modules skip their activation quantiser (disable_network_output_quantization, quant_model.py:93-98;
    # the inner convolutions of the Cheng2020 blocks, quant_block.py:226-300)
    no_act = [n for n, m in qnn.model.named_modules() if isinstance(m, kinds) and m.disable_act_quant]
    blob = dict(format=FORMAT, rules="LU" if lu else "TO", arch=arch, arch_kwargs=dict(arch_kwargs), layers=layers, rest=rest,
                act_bits=first.act_quantizer.n_bits, channel_wise=True, disable_act_quant=no_act)
    if path is not None:
        torch.save(blob, path)
    return blob


def blob_bytes(blob):
    buf = io.BytesIO()
    torch.save(blob, buf)
    return buf.getbuffer().nbytes


def load_int8(blob, device="cuda"):
    """-> QuantModel (TO rules) / light-uniform QuantModel, weights quantised, ready for W-n A-n evaluation."""
    if isinstance(blob, (str, bytes)) or hasattr(blob, "read"):
        blob = torch.load(blob, weights_only=False)
    if blob.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} blob")
    model = codec.ARCHS[blob["arch"]](**blob["arch_kwargs"]).eval()
    missing = model.load_state_dict(blob["rest"], strict=False)
    model.to(device)
    lu = blob["rules"] == "LU"
    bits = {L["n_bits"] for L in blob["layers"].values()}
    wq = dict(n_bits=max(bits), channel_wise=True, scale_method="max")
    if lu:
        qnn = LUQuantModel(model, wq, dict(channel_wise=False, symmetric=False, scale_method="max", leaf_param=True)).eval()
        kinds = (LUQuantModule,)
    else:
        aq = dict(n_bits=blob["act_bits"], channel_wise=True, scale_method="max", leaf_param=False)
        qnn = QuantModel(model, wq, aq, is_cheng=blob["arch"].startswith("cheng")).eval()
        kinds = (QuantModule,)
    mods = {n: m for n, m in qnn.model.named_modules() if isinstance(m, kinds) and getattr(m, "weight", None) is not None}
    if set(mods) != set(blob["layers"]):
        raise ValueError("blob layers do not match the architecture's wrapped layers")
    unexpected = [k for k in missing.unexpected_keys]
    if unexpected:
        raise ValueError(f"unexpected parameters in the blob: {unexpected[:4]}")
    for name, m in mods.items():
        L = blob["layers"][name]
        codes = L["codes"].to(device)
        delta, zp = L["delta"].to(device), L["zero_point"].to(device)
        q = m.weight_quantizer
        q.bitwidth_refactor(L["n_bits"]) if hasattr(q, "bitwidth_refactor") else None
        q.n_bits, q.n_levels = L["n_bits"], 2 ** L["n_bits"]
        shape = ops._bshape(tuple(codes.shape), L["axis"], delta.numel())
        q.delta, q.zero_point, q.inited = delta.view(shape), zp.view(shape), True
        if m.bias is not None:
            m.bias.data.copy_(L["bias"].to(device))
        if lu:
            m.weight.requires_grad_(False)
            m.weight.data = codes.to(torch.uint8)
            m.trained = True
        else:
            if codes.dtype == torch.uint8:
                w = ops.wq_dequant_u8(codes.contiguous(), delta, zp, L["axis"])
            else:                                              # > 8 bit: host-side plumbing of the same expression
                w = (codes.to(torch.float32) - q.zero_point) * q.delta
            m.weight.data.copy_(w)
            m.org_weight = w.clone()
            m.trained = True
    if not lu:
        for m in qnn.model.modules():
            if hasattr(m, "trained"):
                m.trained = True
    no_act = set(blob.get("disable_act_quant", ()))
    for n, m in qnn.model.named_modules():
        if isinstance(m, kinds):
            m.disable_act_quant = n in no_act
    qnn.set_quant_state(True, False)
    return qnn
