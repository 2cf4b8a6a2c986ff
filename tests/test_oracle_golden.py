"""The CPU oracle (oracle/vsm_oracle.py) against the golden vectors produced
by the REAL reference (oracle/make_golden.py).  fp32: rtol 1e-4 / atol 2e-5
(SURVEY.md §8a tolerance policy); ids / argmax / trajectories identical."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vsm_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load_cfg():
    j = json.load(open(os.path.join(G, "tiny_config.json")))
    return O.VSMConfig(**j["cfg"]), j["weight_seed"]


@pytest.fixture(scope="module")
def tiny():
    cfg, seed = load_cfg()
    return cfg, O.synthetic_state_dict(cfg, seed=seed)


def synth_image(seed, w, h):
    from PIL import Image
    return Image.fromarray(np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")


def close(a, b, rtol=1e-4, atol=2e-5):
    a, b = torch.as_tensor(np.asarray(a)).float(), torch.as_tensor(np.asarray(b)).float()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


@pytest.mark.parametrize("tag", ["a", "b"])
def test_model_forward_golden(tiny, tag):
    cfg, sd = tiny
    g = np.load(os.path.join(G, f"model_forward_{tag}.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    ids = torch.from_numpy(g["input_ids"])
    out = O.model_forward_inference(sd, cfg, O.preprocess_owl(img), O.preprocess_clip(img), ids, (int(g["h"]), int(g["w"])))
    r = int(g["loc_row"])
    close(out["hidden"][0, r], g["hidden_loc"])
    close(out["hidden"][0, -1], g["hidden_last"])
    assert np.array_equal(out["logits"][0].argmax(-1).numpy(), g["logits_argmax"])
    close(out["logits"][0, -1], g["logits_last"], atol=1e-4)
    close(out["feature_map"][0, ::6, ::6, :], g["fmap_sample"])
    close(out["low_res_masks"][0, 0], g["low_res_mask"], atol=1e-4)
    pm = out["pred_masks"][0]
    close(pm[::7, ::7], g["pred_mask_sample"], atol=1e-4)
    assert int(pm.argmax()) == int(g["pred_mask_stats"][3])
    close(out["pred_logits"][0, :, 0], g["pred_logits"], atol=1e-4)
    close(out["pred_boxes"][0], g["pred_boxes"])
    assert int(out["pred_logits"][0, :, 0].argmax()) == int(np.argmax(g["pred_logits"]))


def test_generate_golden(tiny):
    cfg, sd = tiny
    g = np.load(os.path.join(G, "generate_a.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    prompt = torch.from_numpy(g["prompt"])
    ids, hidden, am = O.greedy_generate(sd, cfg, prompt, O.preprocess_clip(img), 100, 2, torch.from_numpy(g["forced"]))
    assert am == list(g["argmax"])
    assert np.array_equal(ids.numpy(), g["output_ids"])
    close(hidden[0, -3], g["hidden_loc"])
    _, _, free = O.greedy_generate(sd, cfg, prompt, O.preprocess_clip(img), 4, -1)
    assert free == list(g["free_argmax"])


from tests.helpers import FakeNLP, StubVSM  # noqa: E402  (the stub the goldens were generated with, oracle/make_golden.py)


@pytest.mark.parametrize("tag", ["stub_3lvl", "stub_default", "stub_weakcue", "stub_mixcue"])
def test_search_trajectory_golden(tag):
    from vstar_b200.noun_chunks import extract_noun_chunks
    g = np.load(os.path.join(G, f"search_{tag}.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    kw = json.loads(str(g["kw"]))
    stub = StubVSM()
    nlp = FakeNLP()
    fs, pl, ok, av, path = O.visual_search(stub, img, "mug", None, int(g["smallest"]),
                                           extract_noun_chunks=lambda t: extract_noun_chunks(t, nlp), **kw)
    assert [s.get("context_cue", "") for s in path] == json.loads(str(g["context_cues"]))
    assert np.array_equal(np.array(stub.calls), g["calls"])
    assert np.array_equal(np.array([s["bbox"] for s in path]), g["trajectory"])
    assert pl == int(g["path_length"]) and int(ok) == int(g["success"])
    assert list(fs["bbox"]) == list(g["final_bbox"])
    assert np.array_equal(fs["detection_result"].numpy(), g["detection_result"])


def test_geometry_edge_cases():
    # /root/reference/visual_search.py:234-253
    assert O.split_4subpatches([0, 0, 100, 200]) == (1, 4)
    assert O.split_4subpatches([0, 0, 200, 100]) == (4, 1)
    assert O.split_4subpatches([0, 0, 100, 199]) == (2, 2)
    subs, ws, hs = O.get_sub_patches([3, 5, 101, 77], 2, 2)
    assert subs == [[3, 5, 50, 38], [53, 5, 51, 38], [3, 43, 50, 39], [53, 43, 51, 39]]
    hm = np.zeros((10, 10, 1), np.float32)
    assert all(float(s) == 0 for s in O.get_subpatch_scores(hm, [0, 0, 10, 10], [[0, 0, 5, 5]]))
    flat = torch.full((4, 4, 1), 2.0)
    assert float(O.normalize_score(flat).abs().max()) == 0.0
