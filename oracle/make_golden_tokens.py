"""Golden vectors for the token-major wrappers: outputs of the REFERENCE's own code, imported unmodified
from /root/reference through oracle/_ref_shim.py -- `QuantModule` over nn.Linear and nn.LayerNorm
(task-oriented-PTQ/quantization/quant_layer.py:38-49,105-134), `ActQuantizer` on 3-D tensors (quantizer.py:81-121), and the
`Mlp` of models/layers.py:35-52 wrapped as `QuantMlp` (quant_block.py:330-350).  TEST INFRASTRUCTURE ONLY; runs only where
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
    return G


def main():
    G = token_vectors()
    torch.save(G, OUT)
    print(f"tokens_ref.pt: {len(G)} groups from the reference's own QuantModule / QuantMlp / ActQuantizer "
          f"({os.path.getsize(OUT) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
