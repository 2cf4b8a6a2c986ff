import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstar_b200 import ops
BF = torch.bfloat16
# the dominant launch of the bench step: gate|up projection of the 32-crop frontier batch (M = 32*320)
M, N, K = 10240, 22016, 4096
a = torch.randn(M, K, device="cuda").to(BF)
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
out = torch.empty(M, N // 2, dtype=BF, device="cuda")
for _ in range(3):
    ops.gemm(a, w, out=out, epilogue=ops.EPI_SWIGLU)
torch.cuda.synchronize()
