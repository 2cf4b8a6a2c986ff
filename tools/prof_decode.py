"""one SEAL VQA decode step under `ncu --profile-from-start off --metrics gpu__time_duration.sum` (per-kernel times of a token)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vstar_b200 import synth
from vstar_b200.config import VSMConfig
from vstar_b200.vqa import VQAEngine, VQAWeights

cfg = VSMConfig()
shapes = synth.vqa_state_dict_shapes(cfg)
w = VQAWeights(cfg, lambda n: synth.synthetic_tensor(n, shapes[n], seed=4321, device="cuda"))
eng = VQAEngine(w, max_tokens=1536)
gen = torch.Generator().manual_seed(0)
image = torch.randn(1, 3, 224, 224, generator=gen).to(torch.bfloat16).cuda()
rng = np.random.default_rng(0)
q = [1] + rng.integers(1000, 30000, 40).tolist() + [-200] + rng.integers(1000, 30000, 400).tolist()
x = eng.build_embeds(q, image, None, [True], None)
T = x.shape[0]
eng.prefill_embeds(x)
eng.append_tokens([5], T)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.append_tokens([7], T + 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("T", T)
