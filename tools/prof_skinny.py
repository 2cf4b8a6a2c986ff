"""decode GEMMs under `ncu --set full -k regex:gemm_skinny`: M = 1 (FMA kernel) and M = 8 (mma.sync kernel) on the gate|up shape"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vstar_b200 import ops

N, K = 22016, 4096
w = (torch.randn(N, K, device="cuda") / 64).to(torch.bfloat16)
for M in (1, 8):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    y = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        ops.gemm(a, w, out=y, epilogue=ops.EPI_SWIGLU)
torch.cuda.synchronize()
