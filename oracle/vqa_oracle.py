"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the SEAL VQA-LLM forward that brackets the visual search
(SURVEY.md §8 rows a18-a20).  Same rules as oracle/vsm_oracle.py: only tests/, smoke() and bench.py's baseline legs
may import it; pinned against outputs of the REAL reference (oracle/make_golden.py -> tests/golden/vqa_*.npz).

Restates
  /root/reference/LLaVA/llava/model/llava_search_arch.py:83-279   encode_images / project_features / splice with
                                                                   <image> (-200) and <object> (-300) placeholders
  /root/reference/LLaVA/llava/model/multimodal_projector/builder.py:54-68, perceiver.py:25-121   (object projector)
  /root/reference/LLaVA/llava/model/language_model/llava_search_llama.py:56-141                  (LM forward)
  /root/reference/vstar_bench_eval.py:78-165                       free_form_inference / multiple_choices_inference
State-dict keys = the reference model's own (`model.mm_projector.*`, `model.mm_projector_object.{0,1,2}.*`, ...).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import vsm_oracle as O
from vstar_b200.config import VSMConfig

IMAGE_TOKEN_INDEX = -200
OBJECT_TOKEN_INDEX = -300


def perceiver_resampler(sd, p, x, heads=16, dim_head=96, depth=6):
    """PerceiverResampler.forward (perceiver.py:79-121) for x [B, n, C] -> [B, 32, C]"""
    B, n, C = x.shape
    x = x.unsqueeze(1)                                            # b 1 n d
    x = x + sd[p + "media_pos_emb"][:1]
    lat = sd[p + "latents"].unsqueeze(0).unsqueeze(0).expand(B, 1, -1, -1)
    scale = dim_head ** -0.5
    for i in range(depth):
        a = f"{p}layers.{i}.0."
        xm = F.layer_norm(x, (C,), sd[a + "norm_media.weight"], sd[a + "norm_media.bias"])
        lt = F.layer_norm(lat, (C,), sd[a + "norm_latents.weight"], sd[a + "norm_latents.bias"])
        q = F.linear(lt, sd[a + "to_q.weight"])
        kv = F.linear(torch.cat((xm, lt), dim=-2), sd[a + "to_kv.weight"])
        k, v = kv.chunk(2, dim=-1)

        def sp(t):
            b, m, nn_, _ = t.shape
            return t.view(b, m, nn_, heads, dim_head).permute(0, 3, 1, 2, 4)     # b h t n d

        q, k, v = sp(q), sp(k), sp(v)
        q = q * scale
        sim = torch.einsum("...id,...jd->...ij", q, k)
        sim = sim - sim.amax(dim=-1, keepdim=True)
        attn = sim.softmax(dim=-1)
        out = torch.einsum("...ij,...jd->...id", attn, v)
        out = out.permute(0, 2, 3, 1, 4).reshape(B, 1, lat.shape[2], heads * dim_head)
        lat = F.linear(out, sd[a + "to_out.weight"]) + lat
        f = f"{p}layers.{i}.1."
        h = F.layer_norm(lat, (C,), sd[f + "0.weight"], sd[f + "0.bias"])
        h = F.linear(F.gelu(F.linear(h, sd[f + "1.weight"])), sd[f + "3.weight"])
        lat = h + lat
    res = F.layer_norm(lat, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"])
    return res.squeeze(1)


def project_both(sd, cfg: VSMConfig, pixels):
    """encode_images / project_features: CLIP hidden_states[-2][:,1:] -> (long [B,256,d], short [B,32,d])"""
    feats = O.clip_features(sd, cfg, pixels)
    long_ = F.linear(feats, sd["model.mm_projector.weight"], sd["model.mm_projector.bias"])
    p = "model.mm_projector_object."
    h = F.layer_norm(feats, (feats.shape[-1],), sd[p + "0.weight"], sd[p + "0.bias"])
    h = perceiver_resampler(sd, p + "1.", h)
    short = F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"])
    return long_, short


def splice(sd, ids_1d, img_long, img_short, obj_long, obj_short, images_long, objects_long):
    """prepare_inputs_labels_for_multimodal for one sample (llava_search_arch.py:139-216): one <image>, k <object>s"""
    emb = sd["model.embed_tokens.weight"]
    parts = []
    cur = ids_1d
    pos = (cur == IMAGE_TOKEN_INDEX).nonzero().flatten()
    assert pos.numel() == 1
    s = int(pos[0])
    use_long = images_long is None or bool(images_long[0])
    parts += [emb[cur[:s]], img_long[0] if use_long else img_short[0]]
    cur = cur[s + 1:]
    oi = 0
    while True:
        pos = (cur == OBJECT_TOKEN_INDEX).nonzero().flatten()
        if pos.numel() == 0:
            break
        s = int(pos[0])
        short = objects_long is None or not bool(objects_long[oi])
        parts += [emb[cur[:s]], obj_short[oi] if short else obj_long[oi]]
        oi += 1
        cur = cur[s + 1:]
    if cur.numel() > 0:
        parts.append(emb[cur])
    return torch.cat([t.to(emb.dtype) for t in parts], dim=0)


def build_embeds(sd, cfg, input_ids, image, object_crops=None, images_long=None, objects_long=None):
    img_long, img_short = project_both(sd, cfg, image)
    if object_crops is not None and len(object_crops) > 0:
        obj_long, obj_short = project_both(sd, cfg, object_crops)
    else:
        obj_long = obj_short = None
    return splice(sd, input_ids[0], img_long, img_short, obj_long, obj_short, images_long, objects_long).unsqueeze(0)


def forward_logits(sd, cfg, embeds):
    hidden = O.llama_forward(sd, cfg, embeds)
    return F.linear(hidden, sd["lm_head.weight"])


def free_form_generate(sd, cfg, input_ids, image, object_crops=None, images_long=None, objects_long=None, max_new_tokens=200,
                       eos_token_id=2):
    """model.generate(do_sample=False, use_cache=True) restated as full recompute (mathematically the same greedy ids)"""
    emb = sd["model.embed_tokens.weight"]
    embeds = build_embeds(sd, cfg, input_ids, image, object_crops, images_long, objects_long)
    out = []
    for _ in range(max_new_tokens):
        logits = forward_logits(sd, cfg, embeds)
        nxt = int(torch.argmax(logits[0, -1].float()))
        out.append(nxt)
        if nxt == eos_token_id:
            break
        embeds = torch.cat([embeds, emb[torch.tensor([nxt])].unsqueeze(0)], dim=1)
    return out


def option_losses(sd, cfg, question_ids, options_ids, image, object_crops=None, images_long=None, objects_long=None):
    """multiple_choices_inference (vstar_bench_eval.py:116-165): mean CE of each option's tokens given the question
    (prefix KV shared in the reference; recomputed here).  Returns (losses tensor, argmin)."""
    emb = sd["model.embed_tokens.weight"]
    q_embeds = build_embeds(sd, cfg, question_ids, image, object_crops, images_long, objects_long)
    losses = []
    for opt in options_ids:
        full = torch.cat([q_embeds, emb[opt].unsqueeze(0)], dim=1)
        logits = forward_logits(sd, cfg, full)
        Tq = q_embeds.shape[1]
        lg = logits[0, Tq - 1:Tq - 1 + opt.numel()]              # row predicting each option token
        losses.append(F.cross_entropy(lg.float(), opt))
    losses = torch.stack(losses)
    return losses, int(losses.argmin())


def llama_forward_cached(sd, cfg, embeds, past=None):
    """HF LlamaModel with `use_cache=True` (as driven by vstar_bench_eval.py:127-152): `embeds` [1,Tn,d] are appended after
    the `past` keys/values (list of (k, v) per layer, [1,H,Tp,hd]); -> (final-normed states of the new rows, new past).
    Same arithmetic as vsm_oracle.llama_forward, restricted to the new rows."""
    B, Tn, d = embeds.shape
    H, hd = cfg.n_heads, cfg.head_dim
    Tp = 0 if past is None else past[0][0].shape[2]
    cos, sin = O.rope_cos_sin(Tn, hd, cfg.rope_theta, embeds.dtype, start=Tp)
    mask = torch.full((Tn, Tp + Tn), float("-inf"), dtype=embeds.dtype).triu(Tp + 1)
    x = embeds
    new_past = []
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        h = O.rmsnorm(x, sd[p + "input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, Tn, H, hd).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, Tn, H, hd).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, Tn, H, hd).transpose(1, 2)
        q = q * cos + O.rotate_half(q) * sin
        k = k * cos + O.rotate_half(k) * sin
        if past is not None:
            k, v = torch.cat([past[i][0], k], dim=2), torch.cat([past[i][1], v], dim=2)
        new_past.append((k, v))
        att = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5) + mask
        att = torch.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(att, v).transpose(1, 2).reshape(B, Tn, d)
        x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h = O.rmsnorm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        g = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(g, sd[p + "mlp.down_proj.weight"])
    return O.rmsnorm(x, sd["model.norm.weight"], cfg.rms_eps), new_past


def option_losses_cached(sd, cfg, question_ids, options_ids, image, object_crops=None, images_long=None, objects_long=None):
    """multiple_choices_inference AS WRITTEN (vstar_bench_eval.py:127-163): ONE forward over the question (logits for every
    position, `use_cache=True`), then each option is appended on top of the question's past_key_values.  Same numbers as
    option_losses (which recomputes the prefix per option); this is the variant whose COST equals the reference's."""
    emb = sd["model.embed_tokens.weight"]
    q_embeds = build_embeds(sd, cfg, question_ids, image, object_crops, images_long, objects_long)
    hq, past = llama_forward_cached(sd, cfg, q_embeds)
    q_logits = F.linear(hq, sd["lm_head.weight"])                       # the reference computes them for all T rows
    losses = []
    for opt in options_ids:
        ho, _ = llama_forward_cached(sd, cfg, emb[opt].unsqueeze(0), past)
        lo = F.linear(ho, sd["lm_head.weight"])
        lg = torch.cat([q_logits[0, -1:], lo[0, :-1]], dim=0)            # vstar_bench_eval.py:153
        losses.append(F.cross_entropy(lg.float(), opt))
    losses = torch.stack(losses)
    return losses, int(losses.argmin())


# ---------------------------------------------------------------- host helpers (vstar_bench_eval.py:25-76)
def expand2square_center(pil_img, background_color):
    from PIL import Image
    width, height = pil_img.size
    if width == height:
        return pil_img, 0, 0
    elif width > height:
        result = Image.new(pil_img.mode, (width, width), background_color)
        result.paste(pil_img, (0, (width - height) // 2))
        return result, 0, (width - height) // 2
    else:
        result = Image.new(pil_img.mode, (height, height), background_color)
        result.paste(pil_img, ((height - width) // 2, 0))
        return result, (height - width) // 2, 0


def get_patch(bbox, image_width, image_height, patch_size=224, patch_scale=None):
    import numpy as np
    object_width = int(np.ceil(bbox[2]))
    object_height = int(np.ceil(bbox[3]))
    cx = int(bbox[0] + bbox[2] / 2)
    cy = int(bbox[1] + bbox[3] / 2)
    if patch_scale is None:
        pw, ph = max(object_width, patch_size), max(object_height, patch_size)
    else:
        pw, ph = int(object_width * patch_scale), int(object_height * patch_scale)
    left = max(0, cx - pw // 2)
    right = min(left + pw, image_width)
    top = max(0, cy - ph // 2)
    bottom = min(top + ph, image_height)
    return [left, top, right, bottom]


def vqa_state_dict_shapes(cfg: VSMConfig):
    """tensors of seal_vqa_7b that the path reads (reference key layout); the table itself lives with the synthetic-weight
    helpers (vstar_b200/synth.py) so tools can build a random-init model without importing the oracle"""
    from vstar_b200 import synth
    return synth.vqa_state_dict_shapes(cfg)
