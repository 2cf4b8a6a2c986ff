import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstar_b200 import ops
B, H, S, D = 8, 12, 2305, 64
qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, H * D, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.attn_fused_qkv(qkv, B, S, H, D, False, D ** -0.5, out=out)
torch.cuda.synchronize()
