import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstar_b200 import ops, _lib
from tools.bench_kernels import timeit
BF = torch.bfloat16
for M, N, K, name in [(10240, 22016, 4096, "gate_up B=32"), (10240, 12288, 4096, "qkv B=32"), (10240, 4096, 11008, "down B=32"), (10240, 4096, 4096, "o B=32"),
                      (2560, 22016, 4096, "gate_up B=8"), (2560, 4096, 11008, "down B=8"), (73760, 3072, 768, "owl fc1 B=32"), (73760, 768, 3072, "owl fc2 B=32")]:
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    r = {}
    for gm in (4, 8, 12, 16, 20, 40):
        _lib.call("vsb_gemm_set_group_m", gm)
        r[gm] = round(2 * M * N * K / timeit(lambda: ops.gemm(a, w, out=out), iters=8) / 1e9)
    _lib.call("vsb_gemm_set_group_m", 0)
    print(name, r, flush=True)
