import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstar_b200 import ops, _lib
from tools.bench_kernels import timeit
BF = torch.bfloat16
M, N = 18440, 3072
for K in (768, 1536, 3072):
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
    b = torch.randn(N, device="cuda").to(BF)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    res = {}
    for bn in (256, 512):
        _lib.call("vsb_gemm_set_tuning", bn, 0)
        for tag, kw in (("plain", {}), ("bias", dict(bias=b)), ("qgelu", dict(bias=b, epilogue=ops.EPI_QUICK_GELU)), ("resid", dict(bias=b, residual=out))):
            res[f"{bn}_{tag}"] = round(2 * M * N * K / timeit(lambda: ops.gemm(a, w, out=out, **kw), iters=8) / 1e9)
    _lib.call("vsb_gemm_set_tuning", 0, 0)
    res["cublas"] = round(2 * M * N * K / timeit(lambda: torch.matmul(a, w.t(), out=out), iters=8) / 1e9)
    print(K, res, flush=True)
