"""Full-WIDTH parity (real 7B / CLIP-L / OWL-B dimensions, 2 layers each so the fp32 CPU oracle finishes in seconds):
hidden 4096, 32 heads x 128, intermediate 11008 (SwiGLU interleave, N = 22016), vocab 32004 (lm_head N not a multiple
of 32), CLIP 1024 / 16 heads / 4096, OWL 768 / 12 heads / 3072 / 2305 tokens, query dim 512.  Same tolerance policy as
tests/test_engine_gpu.py (err_new <= 2 * err_ref(bf16 oracle) + floor)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_fullwidth_two_layer_model():
    from PIL import Image
    from oracle import vsm_oracle as O
    from vstar_b200 import ops
    from vstar_b200.config import VSMConfig
    from vstar_b200.engine import VSMEngine, VSMWeights
    torch.set_num_threads(min(32, os.cpu_count() or 8))
    cfg = VSMConfig(n_layers=2, clip_layers=3, owl_layers=2)
    sd = O.synthetic_state_dict(cfg, seed=99)
    eng = VSMEngine(VSMWeights.from_state_dict(cfg, sd), max_tokens=384)
    img = Image.fromarray(np.random.default_rng(7).integers(0, 256, (300, 420, 3), dtype=np.uint8), "RGB")
    prompt, ans = O.synthetic_prompt(cfg, n_text=60, seed=2)
    ids = torch.cat([prompt, ans.unsqueeze(0)], dim=1)
    images, images_clip = O.preprocess_owl(img), O.preprocess_clip(img)
    with torch.no_grad():
        o32 = O.model_forward_inference(sd, cfg, images, images_clip, ids, (300, 420))
        sd16 = {k: v.to(BF) for k, v in sd.items()}
        o16 = O.model_forward_inference(sd16, cfg, images.to(BF), images_clip.to(BF), ids, (300, 420))
    out = eng.model_forward(images.to(BF).cuda(), images_clip.to(BF).cuda(), ids.cuda())
    torch.cuda.synchronize()
    r = int((ids[0] == cfg.loc_token_idx).nonzero()[0, 0]) - 1 + 255
    rep = {}
    for name, new, ref16, ref32, floor in [
        ("hidden_loc", out["hidden_loc"][0], o16["hidden"][0, r], o32["hidden"][0, r], 5e-3),
        ("seg_query", out["seg_queries"][0], o16["seg_queries"][0], o32["seg_queries"][0], 5e-3),
        ("det_query", out["det_queries"][0], o16["det_queries"][0], o32["det_queries"][0], 5e-3),
        ("feature_map", out["feature_map"].view(48, 48, -1), o16["feature_map"][0], o32["feature_map"][0], 1e-2),
        ("low_res", out["low_res_masks"][0], o16["low_res_masks"][0, 0], o32["low_res_masks"][0, 0], 1e-2),
        ("pred_logits", out["pred_logits"][0], o16["pred_logits"][0, :, 0], o32["pred_logits"][0, :, 0], 1e-2),
        ("pred_boxes", out["pred_boxes"][0], o16["pred_boxes"][0], o32["pred_boxes"][0], 1e-2),
    ]:
        e_new, e_ref = rel(new, ref32), rel(ref16, ref32)
        rep[name] = (e_new, e_ref)
        assert e_new <= 2 * e_ref + floor, (name, e_new, e_ref)
    # lm_head over the 5 answer-predicting rows with the odd vocabulary size
    T = ids.shape[1] - 1 + 256
    x, T2, img_pos = eng.prefill(ids.cuda(), images_clip.to(BF).cuda())
    assert T2 == T
    rows = torch.tensor([eng.x_row(0, t) for t in range(T - 6, T)], device="cuda")
    hn, am, logits = eng._logits_rows(x, rows)
    ref_logits = o32["logits"][0, T - 6:T]
    assert logits.shape == ref_logits.shape
    assert rel(logits, ref_logits) <= 2 * rel(o16["logits"][0, T - 6:T], ref_logits) + 1e-2
    top2 = ref_logits.topk(2, -1).values
    conf = (top2[:, 0] - top2[:, 1]) > 0.1
    assert torch.equal(am.cpu().long()[conf], ref_logits.argmax(-1)[conf])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({k: dict(err_new=v[0], err_ref_bf16=v[1]) for k, v in rep.items()}, open("gpurun_out/fullwidth_parity_report.json", "w"), indent=1)
