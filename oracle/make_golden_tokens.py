"""Golden vectors for the token-major wrappers: outputs of the REFERENCE's own code, imported unmodified
from /root/reference through oracle/_ref_shim.py -- `QuantModule` over nn.Linear and nn.LayerNorm
(task-oriented-PTQ/quantization/quant_layer.py:38-49,105-134), `ActQuantizer` on 3-D tensors (quantizer.py:81-121), and the
`Mlp` of models/layers.py:35-52 wrapped as `QuantMlp` (quant_block.py:330-350), the attention core of `WindowAttention`
(models/layers.py:137-166) and an `RSTB` wrapped as `QuantRSTB` (quant_block.py:601-641).  TEST INFRASTRUCTURE ONLY; runs only where
/root/reference exists.  Writes tests/golden/tokens_ref.pt (committed):

    python -m oracle.make_golden_tokens
"""
import copy
import os
import warnings

import torch
import torch.nn as nn

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tokens_ref.pt")
WQ8 = dict(n_bits=8, channel_wise=True, scale_method="max")
AQ8 = dict(n_bits=8, channel_wise=True, scale_method="max", leaf_param=False)


def token_vectors():
    from oracle import _ref_shim as S
    warnings.filterwarnings("ignore")
    S.import_task_oriented()
    from quantization import quant_layer as r_ql, quant_block as r_qb, quantizer as r_q
    from models import layers as r_layers
    g = torch.Generator().manual_seed(1005)
    G = {}
    x = torch.randn(3, 49, 40, generator=g) * 1.5 + 0.2
    cases = {"linear": nn.Linear(40, 100), "linear_nobias": nn.Linear(40, 72, bias=False), "layernorm": nn.LayerNorm(40)}
    with torch.no_grad():
        cases["layernorm"].weight.copy_(1 + 0.3 * torch.randn(40, generator=g))
        cases["layernorm"].bias.copy_(0.1 * torch.randn(40, generator=g))
        for name, mod in cases.items():
            for wbits in (8, 4):
                qm = r_ql.QuantModule(copy.deepcopy(mod), dict(WQ8, n_bits=wbits), AQ8).eval()
                rec = {"x": x.clone(), "state": {k: v.clone() for k, v in mod.state_dict().items()}}
                rec["fp"] = qm(x).clone()
                qm.set_quant_state(True, False)
                rec["w"] = qm(x).clone()
                wq = qm.weight_quantizer
                rec["delta"], rec["zero_point"] = wq.delta.clone(), wq.zero_point.clone()
                rec["w_dq"] = wq(qm.weight).clone()
                qm.trained = True
                qm.set_quant_state(True, True)
                rec["wa"] = qm(x).clone()
                G[f"{name}_w{wbits}"] = rec
        # ActQuantizer on token tensors (per last-axis channel) and the exact GELU
        for i, shape in enumerate([(3, 49, 40), (1, 7, 5), (2, 130, 33)]):
            t = torch.randn(*shape, generator=g) * (1 + i)
            t[..., 0] = 0.25                                   # a constant channel: range clamps to 1e-6
            G[f"actq3d_{i}"] = {"x": t.clone(), "y": r_q.ActQuantizer(t).clone()}
        t = torch.randn(4, 33, 17, generator=g) * 3
        G["gelu"] = {"x": t.clone(), "y": nn.GELU()(t).clone()}
        # QuantMlp (fc1 -> GELU -> [A8] -> fc2), W8 and W8A8
        mlp = r_layers.Mlp(in_features=40, hidden_features=96).eval()
        qmlp = r_qb.QuantMlp(copy.deepcopy(mlp), WQ8, AQ8).eval()
        rec = {"x": x.clone(), "state": {k: v.clone() for k, v in mlp.state_dict().items()}}
        rec["fp"] = qmlp(x).clone()
        qmlp.set_quant_state(True, False)
        rec["w"] = qmlp(x).clone()
        for m in qmlp.modules():
            if hasattr(m, "trained"):
                m.trained = True
        qmlp.set_quant_state(True, True)
        rec["wa"] = qmlp(x).clone()
        G["mlp_w8"] = rec
        # window attention core, from the reference's WindowAttention (models/layers.py:137-166) on shifted windows
        att = r_layers.WindowAttention(32, (4, 4), 4).eval()
        att.relative_position_bias_table.copy_(0.5 * torch.randn(att.relative_position_bias_table.shape, generator=g))
        blk = r_layers.SwinTransformerBlock(32, (8, 12), 4, window_size=4, shift_size=2)
        xw = torch.randn(2 * 6, 16, 32, generator=g)
        qkv = att.qkv(xw)
        B_, N, C = xw.shape
        q, k, v = qkv.reshape(B_, N, 3, 4, C // 4).permute(2, 0, 3, 1, 4)
        a_ = (q * att.scale) @ k.transpose(-2, -1)
        bias = att.relative_position_bias_table[att.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1)
        a_ = a_ + bias.unsqueeze(0)
        a_ = (a_.view(B_ // 6, 6, 4, N, N) + blk.attn_mask.unsqueeze(1).unsqueeze(0)).view(-1, 4, N, N).softmax(-1)
        G["attn_core"] = {"qkv": qkv.clone(), "bias": bias.contiguous().clone(), "mask": blk.attn_mask.clone(),
                          "scale": att.scale, "attn": a_.clone(), "out": (a_ @ v).transpose(1, 2).reshape(B_, N, C).clone(),
                          "x": xw.clone(), "state": {k_: v_.clone() for k_, v_ in att.state_dict().items()},
                          "module_out": att(xw, mask=blk.attn_mask).clone()}
        # QuantRSTB (quant_block.py:601-641) over an RSTB with one plain and one shifted block
        rstb = r_layers.RSTB(dim=32, input_resolution=(8, 8), depth=2, num_heads=4, window_size=4, mlp_ratio=2.).eval()
        for n_, p_ in rstb.named_parameters():
            if "relative_position_bias_table" in n_:
                p_.copy_(0.5 * torch.randn(p_.shape, generator=g))
            elif n_.endswith("norm1.weight") or n_.endswith("norm2.weight"):
                p_.copy_(1 + 0.2 * torch.randn(p_.shape, generator=g))
            elif n_.endswith("bias"):
                p_.copy_(0.1 * torch.randn(p_.shape, generator=g))
            else:
                p_.copy_(0.15 * torch.randn(p_.shape, generator=g))
        xm = torch.randn(2, 32, 8, 8, generator=g)
        qr = r_qb.QuantRSTB(copy.deepcopy(rstb), WQ8, AQ8).eval()
        rec = {"x": xm.clone(), "state": {k_: v_.clone() for k_, v_ in rstb.state_dict().items()},
               "fp_module": rstb(xm, (8, 8)).clone(), "fp": qr(xm, (8, 8)).clone()}
        qr.set_quant_state(True, False)
        rec["w"] = qr(xm, (8, 8)).clone()
        for m in qr.modules():
            if hasattr(m, "trained"):
                m.trained = True
        qr.set_quant_state(True, True)
        rec["wa"] = qr(xm, (8, 8)).clone()
        G["rstb_w8"] = rec
    return G


def main():
    G = token_vectors()
    torch.save(G, OUT)
    print(f"tokens_ref.pt: {len(G)} groups from the reference's own QuantModule / QuantMlp / ActQuantizer "
          f"({os.path.getsize(OUT) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
