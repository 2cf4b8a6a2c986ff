"""Random-init weights / synthetic prompts of the reference architecture (there are no checkpoints
offline; BASELINE.json asks for synthetic data + random-init weights).

Each tensor is drawn from its own generator seeded by (seed, crc32(name)) so any subset can be
regenerated independently: the engine streams a 7B model to the GPU tensor by tensor without ever
materialising a 27 GB fp32 state dict.  CPU generation is what the committed golden vectors use
(tests/golden, oracle/make_golden.py); `device="cuda"` draws from the CUDA generator instead (different
stream, same distribution) for the full-size bench where CPU generation would take minutes.
"""
import math
import zlib

import numpy as np
import torch

from .config import VSMConfig, IMAGE_TOKEN_INDEX


def state_dict_shapes(cfg: VSMConfig):
    """Name -> shape of every tensor the hot path reads (reference key layout)."""
    s = {}
    d, I, V = cfg.hidden, cfg.intermediate, cfg.vocab
    s["model.embed_tokens.weight"] = (V, d)
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        for n in "qkvo":
            s[p + f"self_attn.{n}_proj.weight"] = (d, d)
        s[p + "mlp.gate_proj.weight"] = (I, d)
        s[p + "mlp.up_proj.weight"] = (I, d)
        s[p + "mlp.down_proj.weight"] = (d, I)
        s[p + "input_layernorm.weight"] = (d,)
        s[p + "post_attention_layernorm.weight"] = (d,)
    s["model.norm.weight"] = (d,)
    s["lm_head.weight"] = (V, d)

    def vit(p, C, L, inter, patch, ntok, pre):
        s[p + "embeddings.class_embedding"] = (C,)
        s[p + "embeddings.patch_embedding.weight"] = (C, 3, patch, patch)
        s[p + "embeddings.position_embedding.weight"] = (ntok, C)
        for nm in (pre, "post_layernorm"):
            s[p + nm + ".weight"] = (C,)
            s[p + nm + ".bias"] = (C,)
        for i in range(L):
            q = f"{p}encoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                s[q + f"self_attn.{n}.weight"] = (C, C)
                s[q + f"self_attn.{n}.bias"] = (C,)
            for n in ("layer_norm1", "layer_norm2"):
                s[q + n + ".weight"] = (C,)
                s[q + n + ".bias"] = (C,)
            s[q + "mlp.fc1.weight"] = (inter, C)
            s[q + "mlp.fc1.bias"] = (inter,)
            s[q + "mlp.fc2.weight"] = (C, inter)
            s[q + "mlp.fc2.bias"] = (C,)

    vit("model.vision_tower.vision_tower.vision_model.", cfg.clip_hidden, cfg.clip_layers, cfg.clip_inter,
        cfg.clip_patch, cfg.clip_tokens + 1, "pre_layrnorm")
    s["model.mm_projector.weight"] = (d, cfg.clip_hidden)
    s["model.mm_projector.bias"] = (d,)
    C = cfg.owl_hidden
    vit("model.owlvit.vision_model.", C, cfg.owl_layers, cfg.owl_inter, cfg.owl_patch, cfg.owl_grid ** 2 + 1, "pre_layernorm")
    Q = cfg.owl_query_dim
    s["model.owlvit.class_head.dense0.weight"] = (Q, C); s["model.owlvit.class_head.dense0.bias"] = (Q,)
    for n in ("logit_shift", "logit_scale"):
        s[f"model.owlvit.class_head.{n}.weight"] = (1, C); s[f"model.owlvit.class_head.{n}.bias"] = (1,)
    for n, o in (("dense0", C), ("dense1", C), ("dense2", 4)):
        s[f"model.owlvit.box_head.{n}.weight"] = (o, C); s[f"model.owlvit.box_head.{n}.bias"] = (o,)
    s["model.owlvit.layer_norm.weight"] = (C,); s["model.owlvit.layer_norm.bias"] = (C,)
    D = cfg.sam_dim
    s["model.visual_projection.weight"] = (D, C)
    s["model.prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"] = (2, D // 2)
    s["model.prompt_encoder.no_mask_embed.weight"] = (1, D)
    t = "model.mask_decoder.transformer."

    def att(p, internal):
        for n in ("q_proj", "k_proj", "v_proj"):
            s[p + n + ".weight"] = (internal, D); s[p + n + ".bias"] = (internal,)
        s[p + "out_proj.weight"] = (D, internal); s[p + "out_proj.bias"] = (D,)

    for i in range(cfg.sam_depth):
        lp = f"{t}layers.{i}."
        att(lp + "self_attn.", D)
        att(lp + "cross_attn_token_to_image.", D // 2)
        att(lp + "cross_attn_image_to_token.", D // 2)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            s[lp + n + ".weight"] = (D,); s[lp + n + ".bias"] = (D,)
        s[lp + "mlp.lin1.weight"] = (cfg.sam_mlp, D); s[lp + "mlp.lin1.bias"] = (cfg.sam_mlp,)
        s[lp + "mlp.lin2.weight"] = (D, cfg.sam_mlp); s[lp + "mlp.lin2.bias"] = (D,)
    att(t + "final_attn_token_to_image.", D // 2)
    s[t + "norm_final_attn.weight"] = (D,); s[t + "norm_final_attn.bias"] = (D,)
    m = "model.mask_decoder."
    s[m + "iou_token.weight"] = (1, D); s[m + "mask_tokens.weight"] = (4, D)
    s[m + "output_upscaling.0.conv.weight"] = (D // 4, D, 3, 3); s[m + "output_upscaling.0.conv.bias"] = (D // 4,)
    s[m + "output_upscaling.1.weight"] = (D // 4,); s[m + "output_upscaling.1.bias"] = (D // 4,)
    s[m + "output_upscaling.3.conv.weight"] = (D // 8, D // 4, 3, 3); s[m + "output_upscaling.3.conv.bias"] = (D // 8,)
    for i in range(4):
        for j, o in enumerate((D, D, D // 8)):
            s[f"{m}output_hypernetworks_mlps.{i}.layers.{j}.weight"] = (o, D)
            s[f"{m}output_hypernetworks_mlps.{i}.layers.{j}.bias"] = (o,)
    for w, o in (("det", Q), ("seg", D)):
        s[f"model.text_hidden_fcs_{w}.0.0.weight"] = (d, d); s[f"model.text_hidden_fcs_{w}.0.0.bias"] = (d,)
        s[f"model.text_hidden_fcs_{w}.0.2.weight"] = (o, d); s[f"model.text_hidden_fcs_{w}.0.2.bias"] = (o,)
    return s


def vqa_state_dict_shapes(cfg: VSMConfig):
    """tensors of seal_vqa_7b that the SEAL VQA path reads (reference key layout: LLaVA/llava/model/multimodal_projector/
    builder.py:54-68, perceiver.py:25-121; the Llama / CLIP / linear-projector keys are shared with the VSM table)"""
    s = {k: v for k, v in state_dict_shapes(cfg).items()
         if k.startswith("model.layers.") or k.startswith("model.vision_tower.") or k in
         ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight", "model.mm_projector.weight", "model.mm_projector.bias")}
    C, d = cfg.clip_hidden, cfg.hidden
    p = "model.mm_projector_object."
    s[p + "0.weight"] = (C,); s[p + "0.bias"] = (C,)
    s[p + "1.latents"] = (32, C); s[p + "1.media_pos_emb"] = (1, 1, C)
    inner = 16 * 96
    for i in range(6):
        a = f"{p}1.layers.{i}.0."
        for n in ("norm_media", "norm_latents"):
            s[a + n + ".weight"] = (C,); s[a + n + ".bias"] = (C,)
        s[a + "to_q.weight"] = (inner, C); s[a + "to_kv.weight"] = (2 * inner, C); s[a + "to_out.weight"] = (C, inner)
        f = f"{p}1.layers.{i}.1."
        s[f + "0.weight"] = (C,); s[f + "0.bias"] = (C,)
        s[f + "1.weight"] = (4 * C, C); s[f + "3.weight"] = (C, 4 * C)
    s[p + "1.norm.weight"] = (C,); s[p + "1.norm.bias"] = (C,)
    s[p + "2.weight"] = (d, C); s[p + "2.bias"] = (d,)
    return s


def synthetic_state_dict(cfg: VSMConfig, seed=1234, dtype=torch.float32, scale=None):
    """Deterministic random weights in the reference key layout.  Each tensor
    is drawn from its own generator seeded by (seed, crc32(name)) so any subset
    can be regenerated independently (the GPU engine streams them layer by
    layer without materialising a 27 GB fp32 copy).  Norm weights ~ 1+0.1 N,
    biases 0.02 N, matrices N(0, 1/sqrt(fan_in)) * 0.7 so activations stay O(1)
    through 32 layers (HF std-0.02 init gives a near-linear model whose
    outputs are dominated by the residual stream; this init exercises the
    non-linearities harder)."""
    out = {}
    for name, shape in state_dict_shapes(cfg).items():
        out[name] = synthetic_tensor(name, shape, seed, dtype)
    return out


def synthetic_tensor(name, shape, seed=1234, dtype=torch.float32, device="cpu"):
    g = torch.Generator(device=device).manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    shape = tuple(shape)
    _randn = torch.randn

    def randn(shape, generator):
        return _randn(shape, generator=generator, device=device)

    if "gaussian_matrix" in name:
        t = randn(shape, generator=g)
    elif name.endswith("norm.weight") or "layernorm.weight" in name or "layrnorm.weight" in name \
            or "layer_norm.weight" in name or "layer_norm1.weight" in name or "layer_norm2.weight" in name \
            or name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("norm3.weight") \
            or name.endswith("norm4.weight") or name.endswith("norm_final_attn.weight") \
            or name.endswith("output_upscaling.1.weight"):
        t = 1.0 + 0.1 * randn(shape, generator=g)
    elif name.endswith(".bias"):
        t = 0.02 * randn(shape, generator=g)
    elif len(shape) == 1:       # class_embedding
        t = 0.5 * randn(shape, generator=g)
    elif "embed" in name or name.endswith("_token.weight") or name.endswith("_tokens.weight"):
        t = 0.5 * randn(shape, generator=g)
    else:
        fan_in = int(np.prod(shape[1:]))
        t = (0.7 / math.sqrt(fan_in)) * randn(shape, generator=g)
    return t.to(dtype)


def synthetic_prompt(cfg: VSMConfig, n_text=60, seed=0, answer=True, im_start_index=None):
    """Fixed-length synthetic prompt: BOS, ids, <im_start>, -200,
    <im_end>, ids; optional forced answer [a, b, LOC, c, EOS].  im_start_index places <im_start> (index 37 =
    BOS + 36 ids, the length of the llava_v1 system prompt + "USER:"); default min(37, n_text // 2) as in the goldens."""
    g = torch.Generator().manual_seed(777 + seed)
    hi = min(cfg.vocab - 24, 31990)   # never draws [LOC] / <im_start> / <im_end>
    ids = torch.randint(3, hi, (n_text,), generator=g)
    ids[0] = 1
    im_start, im_end = cfg.vocab - 2, cfg.vocab - 1
    p = min(37, n_text // 2) if im_start_index is None else int(im_start_index)
    assert 1 <= p <= n_text - 3
    ids[p], ids[p + 1], ids[p + 2] = im_start, IMAGE_TOKEN_INDEX, im_end
    ans = torch.tensor([int(ids[3]), int(ids[4]), cfg.loc_token_idx, int(ids[5]), 2])
    return ids.unsqueeze(0), ans
