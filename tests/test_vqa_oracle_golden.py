"""CPU: the VQA oracle (oracle/vqa_oracle.py) against the goldens produced by the REAL reference
LlavaSearchLlamaForCausalLM (oracle/make_golden.py:case_vqa)."""
import os

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as V, vsm_oracle as O
from vstar_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("sub", ["short_long", "long_short"])
def test_vqa_oracle_golden(sub):
    g = np.load(os.path.join(G, f"vqa_a_{sub}.npz"))
    cfg = O.tiny_config()
    sd = {k: synth.synthetic_tensor(k, shp, seed=4321) for k, shp in V.vqa_state_dict_shapes(cfg).items()}
    gen = torch.Generator().manual_seed(int(g["img_seed"]))
    image = torch.randn(1, 3, 224, 224, generator=gen)
    crops = torch.randn(2, 3, 224, 224, generator=gen)
    q = torch.from_numpy(g["q"])
    il, ol = [bool(x) for x in g["images_long"]], [bool(x) for x in g["objects_long"]]
    logits = V.forward_logits(sd, cfg, V.build_embeds(sd, cfg, q, image, crops, il, ol))
    assert logits.shape[1] == int(g["T"])
    assert torch.allclose(logits[0, -1], torch.from_numpy(g["logits_last"]), rtol=1e-4, atol=1e-4)
    assert np.array_equal(logits[0].argmax(-1).numpy(), g["logits_argmax"])
    opts, o = [], 0
    for n in g["opt_lens"]:
        opts.append(torch.from_numpy(g["opts"][o:o + int(n)]))
        o += int(n)
    losses, choice = V.option_losses(sd, cfg, q, opts, image, crops, il, ol)
    assert torch.allclose(losses, torch.from_numpy(g["option_losses"]), rtol=1e-4, atol=1e-4)
    assert choice == int(np.argmin(g["option_losses"]))
    # the as-written variant (question prefilled once, options appended on its past_key_values) gives the same numbers
    losses_c, choice_c = V.option_losses_cached(sd, cfg, q, opts, image, crops, il, ol)
    assert torch.allclose(losses_c, torch.from_numpy(g["option_losses"]), rtol=1e-4, atol=1e-4) and choice_c == choice
    assert V.free_form_generate(sd, cfg, q, image, crops, il, ol, max_new_tokens=4, eos_token_id=-1) == g["gen"].tolist()


def test_tokenizer_image_object_token():
    from vstar_b200.vqa import build_prompt_v1, tokenizer_image_object_token
    from vstar_b200.vsm import SyntheticTokenizer
    tok = SyntheticTokenizer(O.tiny_config())
    p = build_prompt_v1("<image>\nIs the mug <object> at [0.1,0.2,0.3,0.4] next to the cup <object> ?")
    ids = tokenizer_image_object_token(p, tok)
    assert ids[0] == 1 and ids.count(-200) == 1 and ids.count(-300) == 2 and ids.index(-200) < ids.index(-300)
    full = tokenizer_image_object_token(build_prompt_v1("<image>\nq", "yes"), tok)
    assert full[:len(tokenizer_image_object_token(build_prompt_v1("<image>\nq"), tok))] == tokenizer_image_object_token(build_prompt_v1("<image>\nq"), tok)
