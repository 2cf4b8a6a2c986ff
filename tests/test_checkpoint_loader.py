"""Checkpoint wire format: HF sharded safetensors / .bin with the key layout written by
merge_lora_weights_and_save_hf_model.py:143-151 -> kernel-side weight layouts.  CPU-only (torch views, no kernels)."""
import json
import os

import torch

from vstar_b200 import synth
from vstar_b200.config import tiny_config
from vstar_b200.engine import VSMWeights
from vstar_b200.vsm import config_from_hf, open_checkpoint


def _write_ckpt(tmp, cfg, sd, fmt):
    os.makedirs(tmp, exist_ok=True)
    json.dump(dict(hidden_size=cfg.hidden, num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_heads,
                   intermediate_size=cfg.intermediate, vocab_size=cfg.vocab, rms_norm_eps=cfg.rms_eps, out_dim=cfg.owl_query_dim),
              open(os.path.join(tmp, "config.json"), "w"))
    keys = sorted(sd)
    half = len(keys) // 2
    shards = [{k: sd[k].contiguous() for k in keys[:half]}, {k: sd[k].contiguous() for k in keys[half:]}]
    if fmt == "safetensors":
        from safetensors.torch import save_file
        for i, sh in enumerate(shards):
            save_file(sh, os.path.join(tmp, f"model-{i + 1:05d}-of-00002.safetensors"))
    else:
        for i, sh in enumerate(shards):
            torch.save(sh, os.path.join(tmp, f"pytorch_model-{i + 1:05d}-of-00002.bin"))


def test_sharded_checkpoint_roundtrip(tmp_path):
    cfg = tiny_config()
    sd = synth.synthetic_state_dict(cfg, seed=7)
    clip_pfx = "model.vision_tower.vision_tower."
    main = {k: v.to(torch.bfloat16) for k, v in sd.items() if not k.startswith(clip_pfx)}       # CLIP keys are dropped (merge...py:146-149)
    clip = {k[len(clip_pfx):]: v for k, v in sd.items() if k.startswith(clip_pfx)}
    for fmt in ("safetensors", "bin"):
        d1, d2 = str(tmp_path / f"vsm_{fmt}"), str(tmp_path / f"clip_{fmt}")
        _write_ckpt(d1, cfg, main, fmt)
        _write_ckpt(d2, cfg, clip, fmt)
        c2 = config_from_hf(d1)
        assert (c2.hidden, c2.n_layers, c2.n_heads, c2.intermediate, c2.vocab, c2.owl_query_dim) == \
               (cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.intermediate, cfg.vocab, cfg.owl_query_dim)
        gm, gc = open_checkpoint(d1), open_checkpoint(d2)

        def get(name):
            return gc(name[len(clip_pfx):]) if name.startswith(clip_pfx) else gm(name)

        w = VSMWeights(cfg, get, device="cpu")
        bf = lambda t: t.to(torch.bfloat16)
        L0 = w.layers[0]
        p = "model.layers.0."
        # the RMSNorm weights are folded into the projections that consume the normalised activations (engine.CoreWeights)
        fold = lambda w_, ln: (bf(w_).float() * bf(sd[p + ln]).float()[None, :]).to(torch.bfloat16)
        assert w.fold_norms and torch.equal(L0["ln1"], torch.ones_like(L0["ln1"]))
        assert torch.equal(L0["wqkv"], torch.cat([fold(sd[p + "self_attn.q_proj.weight"], "input_layernorm.weight"),
                                                  fold(sd[p + "self_attn.k_proj.weight"], "input_layernorm.weight"),
                                                  fold(sd[p + "self_attn.v_proj.weight"], "input_layernorm.weight")], 0))
        assert torch.equal(L0["wgu"][0::2], fold(sd[p + "mlp.gate_proj.weight"], "post_attention_layernorm.weight"))
        assert torch.equal(L0["wgu"][1::2], fold(sd[p + "mlp.up_proj.weight"], "post_attention_layernorm.weight"))
        assert torch.equal(w.clip["patch_w"][:, :588], bf(sd[clip_pfx + "vision_model.embeddings.patch_embedding.weight"]).reshape(-1, 588))
        assert w.cls_w.shape == (cfg.owl_query_dim + 2, cfg.owl_hidden)
        conv = bf(sd["model.mask_decoder.output_upscaling.0.conv.weight"])
        assert torch.equal(w.up0_w.view(conv.shape[0], 3, 3, conv.shape[1]), conv.permute(0, 2, 3, 1))
        assert w.dense_pe.shape == (48 * 48, 256) and w.box_bias.shape == (2304, 4)
