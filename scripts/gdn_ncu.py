"""One launch of each variant of the one-kernel GDN forward for an `ncu --set full -k gdn_fused_kernel` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rdo_ptq_b200 import ops
shape = (1, 192, 384, 512)
N, Cc, H, W = shape
x = torch.randn(shape, device="cuda")
gam = torch.rand(Cc, Cc, device="cuda") * 0.02 + 0.1 * torch.eye(Cc, device="cuda")
bet = 1 + torch.rand(Cc, device="cuda")
d = ops.gdn_desc(x.shape, False)
packed = ops.pack_weights(gam.view(Cc, Cc, 1, 1), d, False)
keys = ops.act_quant_stats(x)
y = torch.empty_like(x)
for _ in range(2):
    ops.gdn_fwd_fused(x, packed, bet, False, pending=(keys, 8), y=y)
    ops.gdn_fwd_fused(x, packed, bet, False, y=y)
torch.cuda.synchronize()
