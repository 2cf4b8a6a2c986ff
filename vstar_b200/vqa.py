"""SEAL VQA LLM on the sm_100a kernels — the model that brackets the visual search in `vstar_bench_eval.py`
(SURVEY.md §8 rows a18-a20, §8f-1).

  VQAWeights / VQAEngine : LlavaSearchLlamaForCausalLM (/root/reference/LLaVA/llava/model/language_model/llava_search_llama.py:56-141,
                           llava_search_arch.py:83-279): CLIP-L -> BOTH projectors (linear "long" 256 tokens; LayerNorm ->
                           PerceiverResampler(depth 6, 16 heads x 96, 32 latents) -> Linear "short" 32 tokens), <image>/<object>
                           splicing with the long/short switches, Llama prefill / greedy decode on the fused-QKV cache, option
                           scoring with the question prefix kept in the cache (vstar_bench_eval.py:116-165).
  VQA_LLM                : drop-in mirror of `vstar_bench_eval.VQA_LLM` (free_form_inference, multiple_choices_inference,
                           get_patch, get_object_crop).

The reference runs this model in fp16 (LLaVA/llava/model/builder.py:43); the kernels here are bf16 with fp32
accumulation — the tolerance is stated in tests/test_vqa_gpu.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .config import VSMConfig
from .engine import CoreWeights, LlamaClipCore
from .vsm import CLIP_MEAN, SyntheticTokenizer, _normalize_into

BF = torch.bfloat16
IMAGE_TOKEN_INDEX = -200
OBJECT_TOKEN_INDEX = -300
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_OBJECT_TOKEN = "<object>"

VICUNA_V1_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                    "The assistant gives helpful, detailed, and polite answers to the user's questions.")


def build_prompt_v1(question_with_image_token, answer=None):
    """conv_templates['v1'] (LLaVA/llava/conversation.py:252-262; SeparatorStyle.TWO, sep ' ', sep2 '</s>')"""
    ret = VICUNA_V1_SYSTEM + " " + "USER: " + question_with_image_token + " "
    if answer:
        ret += "ASSISTANT: " + answer + "</s>"
    else:
        ret += "ASSISTANT:"
    return ret


def tokenizer_image_object_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, object_token_index=OBJECT_TOKEN_INDEX):
    """LLaVA/llava/mm_utils.py:65-87 (behaviour kept verbatim: defines the token sequence)"""
    prompt_chunks = []
    for prompt_chunk in prompt.split("<image>"):
        prompt_chunks.extend(prompt_chunk.split("<object>"))
    prompt_chunks = [tokenizer(chunk).input_ids for chunk in prompt_chunks]

    def insert_separator(X, seps):
        return [ele for sublist in zip(X, seps) for ele in sublist][:-1]

    input_ids = []
    offset = 0
    if len(prompt_chunks) > 0 and len(prompt_chunks[0]) > 0 and prompt_chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        input_ids.append(prompt_chunks[0][0])
    sep = [[image_token_index] * (offset + 1)] + [[object_token_index] * (offset + 1)] * (len(prompt_chunks) - 1)
    for x in insert_separator(prompt_chunks, sep):
        input_ids.extend(x[offset:])
    return input_ids


class VQAWeights(CoreWeights):
    """CoreWeights + the object projector `model.mm_projector_object.{0,1,2}` (multimodal_projector/builder.py:54-68)"""

    def __init__(self, cfg: VSMConfig, get, device="cuda"):
        super().__init__(cfg, get, device)
        g = self._g
        p = "model.mm_projector_object."
        self.pj_ln = (g(p + "0.weight"), g(p + "0.bias"))
        self.latents = g(p + "1.latents")
        self.media_pos = g(p + "1.media_pos_emb").reshape(1, -1).contiguous()
        self.pj_layers = []
        for i in range(6):
            a, f = f"{p}1.layers.{i}.0.", f"{p}1.layers.{i}.1."
            self.pj_layers.append(dict(
                nm=(g(a + "norm_media.weight"), g(a + "norm_media.bias")), nl=(g(a + "norm_latents.weight"), g(a + "norm_latents.bias")),
                wq=g(a + "to_q.weight"), wkv=g(a + "to_kv.weight"), wo=g(a + "to_out.weight"),
                ffn=(g(f + "0.weight"), g(f + "0.bias")), w1=g(f + "1.weight"), w2=g(f + "3.weight")))
        self.pj_norm = (g(p + "1.norm.weight"), g(p + "1.norm.bias"))
        self.pj_out = (g(p + "2.weight"), g(p + "2.bias"))


class VQAEngine(LlamaClipCore):
    HEADS, DHEAD, NLAT = 16, 96, 32

    def __init__(self, weights: VQAWeights, max_tokens=1536):
        super().__init__(weights, max_tokens=max_tokens)

    # ------------------------------------------------------------------ projectors
    def project_both(self, pixels):
        """encode_images / project_features (llava_search_arch.py:83-93): pixels [n,3,224,224] bf16 ->
        (long [n*256, d], short [n*32, d])"""
        c, w = self.cfg, self.w
        n = pixels.shape[0]
        ct, S = self.clip_tokens(pixels)                                   # [n*257, C] incl. CLS
        P = S - 1
        rows = torch.arange(n * S, device=self.dev).view(n, S)[:, 1:].reshape(-1).contiguous()
        feats = ops.gather_rows(rows, ct)                                  # hidden_states[-2][:, 1:]
        long_ = ops.gemm(feats, w.mm_w, bias=w.mm_b)
        C = feats.shape[1]
        inner = self.HEADS * self.DHEAD
        x = ops.layernorm(feats, w.pj_ln[0], w.pj_ln[1], 1e-5)
        x = ops.add_rows(x, w.media_pos)                                   # + media_pos_emb[:1]
        lat = w.latents.repeat(n, 1).contiguous()                          # [n*32, C]
        NK = P + self.NLAT
        kv = torch.empty((n * NK, 2 * inner), dtype=BF, device=self.dev)
        for L in w.pj_layers:
            xm = ops.layernorm(x, L["nm"][0], L["nm"][1], 1e-5)
            lt = ops.layernorm(lat, L["nl"][0], L["nl"][1], 1e-5)
            q = ops.gemm(lt, L["wq"])
            # to_kv(cat(media, latents)): two GEMMs scatter their rows into one [n, P+32, 2*inner] buffer
            ops.gemm(xm, L["wkv"], out=kv, rows_per_group=P, group_stride=NK, group_offset=0)
            ops.gemm(lt, L["wkv"], out=kv, rows_per_group=self.NLAT, group_stride=NK, group_offset=P)
            o = ops.attn_small(q, kv[:, :inner], kv[:, inner:], n, self.HEADS, self.NLAT, NK, self.DHEAD, self.DHEAD ** -0.5)
            lat = ops.gemm(o, L["wo"], residual=lat)
            h = ops.layernorm(lat, L["ffn"][0], L["ffn"][1], 1e-5)
            h = ops.gemm(h, L["w1"], epilogue=ops.EPI_GELU)
            lat = ops.gemm(h, L["w2"], residual=lat)
        res = ops.layernorm(lat, w.pj_norm[0], w.pj_norm[1], 1e-5)
        short = ops.gemm(res, w.pj_out[0], bias=w.pj_out[1])
        return long_, short

    # ------------------------------------------------------------------ splice (llava_search_arch.py:139-216)
    def build_embeds(self, input_ids, image, object_crops=None, images_long=None, objects_long=None):
        """input_ids: python list with one -200 and k -300 placeholders -> embeds [T, d] on the device"""
        return self.build_embeds_batch([(input_ids, image, object_crops, images_long, objects_long)])[0]

    def _splice(self, ids, img_long, img_short, obj_long, obj_short, images_long, objects_long):
        c = self.cfg
        segs = []                                   # ("ids", list) | ("feat", tensor [m,d])
        s = ids.index(IMAGE_TOKEN_INDEX)
        use_long = images_long is None or bool(images_long[0])
        segs += [("ids", ids[:s]), ("feat", img_long if use_long else img_short)]
        cur = ids[s + 1:]
        oi = 0
        while OBJECT_TOKEN_INDEX in cur:
            s = cur.index(OBJECT_TOKEN_INDEX)
            short = objects_long is None or not bool(objects_long[oi])
            f = obj_short[oi * 32:(oi + 1) * 32] if short else obj_long[oi * c.clip_tokens:(oi + 1) * c.clip_tokens]
            segs += [("ids", cur[:s]), ("feat", f)]
            oi += 1
            cur = cur[s + 1:]
        if cur:
            segs.append(("ids", cur))
        T = sum(len(v) if k == "ids" else v.shape[0] for k, v in segs)
        x = torch.empty((T, c.hidden), dtype=BF, device=self.dev)
        t = 0
        for k, v in segs:
            n = len(v) if k == "ids" else v.shape[0]
            if n == 0:
                continue
            if k == "ids":
                ops.gather_rows(torch.tensor(v, dtype=torch.int64, device=self.dev), self.w.embed, out=x[t:t + n])
            else:
                ops.copy2d(v, x[t:t + n])
            t += n
        return x

    # ------------------------------------------------------------------ LM
    def prefill_embeds(self, x, reserve=0):
        """x [T,d] (consumed in place) -> residual stream after all layers; cache positions 0..T.  reserve: positions the
        caller will append (the slot is sized for T + reserve; beyond the 2048 RoPE rows raises VsbError)"""
        T = x.shape[0]
        self._ensure_cache(1, self._capacity(T + reserve))
        self._prefix_slots = 0
        self._llm_layers(x, 1, T, 0, self.max_tokens)
        self.kv_epoch = getattr(self, "kv_epoch", 0) + 1        # handles to an older prefix are stale from here on
        return x

    def append_tokens(self, tokens, past):
        """run `tokens` (python list) on top of `past` cached positions -> logits fp32 [n, V] for every new position"""
        ids = torch.tensor(tokens, dtype=torch.int64, device=self.dev)
        x = ops.gather_rows(ids, self.w.embed)
        if past + len(tokens) > self._cache_shape[2]:
            self._ensure_cache(self._cache_shape[1], self._capacity(past + len(tokens) + 64), keep=True)
        self._llm_layers(x, 1, len(tokens), past, self.max_tokens)
        hn = ops.rmsnorm(x, self.w.final_norm, self.cfg.rms_eps)
        return ops.gemm(hn, self.w.lm_head, out_dtype=torch.float32)

    def last_logits(self, x):
        rows = torch.tensor([x.shape[0] - 1], dtype=torch.int64, device=self.dev)
        hn, am, logits = self._logits_rows(x, rows)
        return logits

    def generate(self, input_ids, image, object_crops=None, images_long=None, objects_long=None, max_new_tokens=200, eos_token_id=2,
                 stop_ids=None, stop_fn=None):
        """greedy `model.generate(use_cache=True, do_sample=False)` (vstar_bench_eval.py:91-103) -> list of new token ids.
        stop_ids: stop once the output ends with these ids; stop_fn(new_ids) -> bool: HF-style stopping criterion."""
        x = self.build_embeds(input_ids, image, object_crops, images_long, objects_long)
        T = x.shape[0]
        self.prefill_embeds(x, reserve=max(0, min(max_new_tokens, self.MAX_POSITIONS - T)))
        logits = self.last_logits(x)
        out = []
        past = T
        for _ in range(max_new_tokens):
            nxt = int(ops.argmax_rows(logits)[0][0])
            out.append(nxt)
            if nxt == eos_token_id or (stop_ids and out[-len(stop_ids):] == list(stop_ids)) or \
                    (stop_fn is not None and stop_fn(out)):
                break
            logits = self.append_tokens([nxt], past)
            past += 1
        return out

    # ------------------------------------------------------------------ continuous-batched greedy decode
    def prefill_ragged(self, xs, reserve=0):
        """xs: list of [T_b, d] input embeddings (consumed).  Sequences are LEFT-padded in the shared cache: sequence b
        occupies cache rows [Tpad - T_b, Tpad) of batch slot b, so every sequence ends at the same row and one decode
        step is ONE set of kernels for the whole batch (each weight byte read once for B tokens).  Runs of neighbouring
        sequences of EQUAL length are prefilled as one batch (M = G * T rows) - generate_batch orders its requests by length.
        -> (last-position logits fp32 [B, V], Tpad, lengths)"""
        B = len(xs)
        lens = [int(x.shape[0]) for x in xs]
        Tpad = max(lens)
        self._ensure_cache(B, self._capacity(Tpad + reserve))
        self._prefix_slots = 0
        Tm = self._cache_shape[2]
        last = []
        b = 0
        while b < B:
            G = 1
            while b + G < B and lens[b + G] == lens[b]:
                G += 1
            T = lens[b]
            x = torch.cat(xs[b:b + G], 0).contiguous() if G > 1 else xs[b]
            self._llm_layers(x, G, T, 0, Tm, cache_row_offset=b * Tm + (Tpad - T))
            last += [x[g * T + T - 1:g * T + T] for g in range(G)]
            b += G
        self.kv_epoch = getattr(self, "kv_epoch", 0) + 1
        hn = ops.rmsnorm(torch.cat(last, 0).contiguous(), self.w.final_norm, self.cfg.rms_eps)
        return ops.gemm(hn, self.w.lm_head, out_dtype=torch.float32), Tpad, lens

    def decode_ragged(self, tokens, lens, Tpad, step):
        """one decode step for the left-padded batch: tokens int64 [B] (token `step` of every answer) -> logits fp32 [B, V]"""
        B = len(lens)
        ids = torch.as_tensor(tokens, dtype=torch.int64, device=self.dev)
        x = ops.gather_rows(ids, self.w.embed)
        if Tpad + step + 1 > self._cache_shape[2]:
            self._ensure_cache(self._cache_shape[1], self._capacity(Tpad + step + 64), keep=True)
        positions = torch.tensor([n + step for n in lens], dtype=torch.int32, device=self.dev)
        k_start = torch.tensor([Tpad - n for n in lens], dtype=torch.int32, device=self.dev)
        self._llm_layers(x, B, 1, Tpad + step, self._cache_shape[2], positions=positions, k_start=k_start)
        hn = ops.rmsnorm(x, self.w.final_norm, self.cfg.rms_eps)
        return ops.gemm(hn, self.w.lm_head, out_dtype=torch.float32)

    def generate_batch(self, requests, max_new_tokens=200, eos_token_id=2, stop_ids=None):
        """greedy generation for several (input_ids, image, object_crops, images_long, objects_long) requests at once.
        Same per-sequence semantics as generate(); finished sequences idle (their slots keep stepping, results ignored).
        -> list of new-token lists"""
        xs = self.build_embeds_batch(list(requests))           # one CLIP + projector pass for every image of the batch
        perm = sorted(range(len(xs)), key=lambda i: int(xs[i].shape[0]))       # equal lengths next to each other (stable)
        xs = [xs[i] for i in perm]
        longest = max(int(x.shape[0]) for x in xs)
        logits, Tpad, lens = self.prefill_ragged(xs, reserve=max(0, min(max_new_tokens, self.MAX_POSITIONS - longest)))
        B = len(xs)
        outs = [[] for _ in range(B)]
        done = [False] * B
        for step in range(max_new_tokens):
            nxt = ops.argmax_rows(logits)[0].tolist()              # one D2H per step for the whole batch
            for b in range(B):
                if done[b]:
                    continue
                outs[b].append(int(nxt[b]))
                if nxt[b] == eos_token_id or (stop_ids and outs[b][-len(stop_ids):] == list(stop_ids)):
                    done[b] = True
            if all(done) or step == max_new_tokens - 1:
                break
            logits = self.decode_ragged(nxt, lens, Tpad, step)
        res = [None] * B
        for k, i in enumerate(perm):
            res[i] = outs[k]
        return res

    def option_losses(self, question_ids, options_ids, image, object_crops=None, images_long=None, objects_long=None):
        """multiple_choices_inference core (vstar_bench_eval.py:127-163).  Returns (losses fp32 [n_options] on host, argmin)."""
        return self.option_losses_batch([(question_ids, options_ids, image, object_crops, images_long, objects_long)])[0]

    def build_embeds_batch(self, items):
        """items: [(input_ids, image [1,3,224,224], object_crops [k,3,224,224] | None, images_long, objects_long)] -> list of
        embeds [T_i, d].  ONE CLIP + projector pass over every image and object crop of the batch."""
        c = self.cfg
        pix, spans = [], []
        for ids, image, crops, il, ol in items:
            k = 0 if crops is None else int(crops.shape[0])
            spans.append((len(pix), k))
            pix.append(image)
            if k:
                pix.append(crops)
        allp = torch.cat(pix, 0).contiguous() if len(pix) > 1 else pix[0]
        long_all, short_all = self.project_both(allp)
        nl, ns = c.clip_tokens, self.NLAT
        out, img_i = [], 0
        for (ids, image, crops, il, ol) in items:
            k = 0 if crops is None else int(crops.shape[0])
            img_long, img_short = long_all[img_i * nl:(img_i + 1) * nl], short_all[img_i * ns:(img_i + 1) * ns]
            obj_long, obj_short = long_all[(img_i + 1) * nl:(img_i + 1 + k) * nl], short_all[(img_i + 1) * ns:(img_i + 1 + k) * ns]
            img_i += 1 + k
            out.append(self._splice(list(ids), img_long, img_short, obj_long, obj_short, il, ol))
        return out

    def option_losses_batch(self, items):
        """Option scoring for several questions: items = [(question_ids, options_ids, image, object_crops, images_long,
        objects_long)] -> [(losses fp32 [n_options] on host, argmin)].

        The reference prefills the question once and then runs ONE FORWARD PER OPTION on its past_key_values
        (vstar_bench_eval.py:140-152).  Here every option of a question is scored in ONE pass: the options' tokens are
        appended back to back after the question's cache rows, every row carries its own RoPE position (Tq + index inside its
        option) and the attention masks the other options' rows (vsb_flash_attn_seg_bf16), so each row sees exactly the keys
        it sees in the reference.  Questions whose spliced length is equal are batched: one prefill with M = G*Tq rows and
        one option pass for all of them (each weight byte is read twice per GROUP instead of five times per question)."""
        embeds = self.build_embeds_batch([(q, img, crops, il, ol) for q, opts, img, crops, il, ol in items])
        groups = {}
        for i, x in enumerate(embeds):
            groups.setdefault(int(x.shape[0]), []).append(i)
        results = [None] * len(items)
        dev = self.dev
        for Tq, members in groups.items():
            G = len(members)
            opts = [[list(o) for o in items[i][1]] for i in members]
            Tn = max(sum(len(o) for o in oo) for oo in opts)
            self._ensure_cache(G, self._capacity(Tq + Tn))
            self._prefix_slots = 0
            x = torch.cat([embeds[i] for i in members], 0).contiguous() if G > 1 else embeds[members[0]]
            self._llm_layers(x, G, Tq, 0, self.max_tokens)
            self.kv_epoch = getattr(self, "kv_epoch", 0) + 1
            last = torch.tensor([g * Tq + Tq - 1 for g in range(G)], dtype=torch.int64, device=dev)
            _, _, q_last = self._logits_rows(x, last)                          # [G, V]: predicts the first token of every option
            # appended rows: options back to back, right-padded to Tn with rows that see only themselves
            ids = torch.zeros((G, Tn), dtype=torch.int64)
            pos = torch.zeros((G, Tn), dtype=torch.int32)
            seg = torch.zeros((G, Tn), dtype=torch.int32)
            pred, labels, spans = [], [], []
            for g, oo in enumerate(opts):
                off = 0
                for o in oo:
                    spans.append((g, len(labels), len(o)))
                    for j, tok in enumerate(o):
                        ids[g, off + j], pos[g, off + j], seg[g, off + j] = tok, Tq + j, Tq + off
                        pred.append(g if j == 0 else G + g * Tn + off + j - 1)
                        labels.append(tok)
                    off += len(o)
                for r in range(off, Tn):
                    pos[g, r], seg[g, r] = Tq, Tq + r
            xo = ops.gather_rows(ids.view(-1).to(dev), self.w.embed)
            self._llm_layers(xo, G, Tn, Tq, self.max_tokens, positions=pos.view(-1).to(dev), q_seg=seg.view(-1).to(dev), seg_lo=Tq)
            hn = ops.rmsnorm(xo, self.w.final_norm, self.cfg.rms_eps)
            lo = ops.gemm(hn, self.w.lm_head, out_dtype=torch.float32)
            rows = torch.cat([q_last, lo], 0).index_select(0, torch.tensor(pred, dtype=torch.int64, device=dev)).contiguous()
            nll = ops.nll_rows(rows, torch.tensor(labels, dtype=torch.int64, device=dev)).cpu()
            per = {}
            for g, s0, n in spans:
                per.setdefault(g, []).append(nll[s0:s0 + n].mean())
            for g, i in enumerate(members):
                losses = torch.stack(per[g])
                results[i] = (losses, int(losses.argmin()))
        return results


# ----------------------------------------------------------------------------------------------------------------
def _clip_preprocess(pil_img, size=224):
    """CLIPImageProcessor.preprocess of the reference (shortest edge 224 bicubic, centre crop 224, /255, mean/std)"""
    from PIL import Image
    img = pil_img.convert("RGB")
    w, h = img.size
    short = min(w, h)
    if (w, h) != (size, size):
        nw, nh = int(w * size / short), int(h * size / short)
        img = img.resize((nw, nh), resample=Image.BICUBIC)
        left, top = (nw - size) // 2, (nh - size) // 2
        img = img.crop((left, top, left + size, top + size))
    out = torch.empty((3, size, size), dtype=torch.float32)
    _normalize_into(np.array(img), out)
    return out


class _Proc:
    crop_size = {"height": 224, "width": 224}
    image_mean = list(CLIP_MEAN)

    def preprocess(self, image, return_tensors="pt"):
        return {"pixel_values": [_clip_preprocess(image)]}


class _ModelConfig:
    pass


class _ModelOutput:
    def __init__(self, logits, past_key_values):
        self.logits, self.past_key_values = logits, past_key_values


class _Past:
    """opaque `past_key_values`: the KV rows live in the engine's cache; this records how many are valid and which prefill
    wrote them (a handle from before a later prefill is rejected instead of silently reading the wrong rows)"""

    def __init__(self, length, epoch):
        self.length, self.epoch = length, epoch


class LlavaSearchModel:
    """The `model` that load_pretrained_model returns: the part of LlavaSearchLlamaForCausalLM's surface that
    /root/reference/vstar_bench_eval.py drives (generate(...) at :91-103; forward with images / with past_key_values at
    :127-152; .config.vocab_size at :157), on the sm_100a engine.  Greedy decoding only (the benchmark uses temperature 0)."""

    def __init__(self, engine: VQAEngine, eos_token_id=2):
        self.engine = engine
        self.config = _ModelConfig()
        self.config.vocab_size = engine.cfg.vocab
        self.eos_token_id = eos_token_id
        self.device = torch.device(engine.dev)

    def eval(self):
        return self

    def cuda(self):
        return self

    def _px(self, t):
        if t is None or len(t) == 0:
            return None
        return ops.cast_f32_bf16(torch.as_tensor(t).float().to(self.engine.dev).contiguous())

    @torch.inference_mode()
    def generate(self, input_ids, images=None, object_features=None, images_long=None, objects_long=None, do_sample=False, num_beams=1,
                 temperature=0, top_p=None, max_new_tokens=200, use_cache=True, stopping_criteria=None, **_):
        if do_sample or num_beams != 1:
            raise NotImplementedError("greedy decoding only (vstar_bench_eval.py:196 runs temperature=0, num_beams=1)")
        ids = input_ids.view(-1).tolist()
        stop_fn = None
        if stopping_criteria:
            def stop_fn(new):
                full = torch.tensor([ids + list(new)], dtype=torch.int64)
                return any(bool(sc(full, None)) for sc in stopping_criteria)
        new = self.engine.generate(ids, self._px(images), self._px(object_features), images_long, objects_long, max_new_tokens,
                                   self.eos_token_id, stop_fn=stop_fn)
        return torch.cat([input_ids.view(1, -1).cpu(), torch.tensor([new], dtype=torch.int64)], dim=1).to(input_ids.device)

    @torch.inference_mode()
    def __call__(self, input_ids=None, use_cache=True, images=None, object_features=None, images_long=None, objects_long=None,
                 attention_mask=None, past_key_values=None, **_):
        e = self.engine
        if past_key_values is None:
            x = e.build_embeds(input_ids.view(-1).tolist(), self._px(images), self._px(object_features), images_long, objects_long)
            T = x.shape[0]
            e.prefill_embeds(x)
            _, _, logits = e._logits_rows(x, torch.arange(T, dtype=torch.int64, device=e.dev))
            return _ModelOutput(logits.view(1, T, -1), _Past(T, e.kv_epoch))
        if past_key_values.epoch != e.kv_epoch:
            raise RuntimeError("stale past_key_values: the engine's KV cache has been re-prefilled since this handle was returned")
        toks = input_ids.view(-1).tolist()
        logits = e.append_tokens(toks, past_key_values.length)          # rows of an earlier continuation are overwritten
        return _ModelOutput(logits.view(1, len(toks), -1), _Past(past_key_values.length + len(toks), e.kv_epoch))


def load_pretrained_model(model_path, model_base=None, model_name=None, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda"):
    """Drop-in for /root/reference/LLaVA/llava/model/builder.py:26-151 (the SEAL VQA LLM branch):
    -> (tokenizer, model, image_processor, context_len).  model_path: local `seal_vqa_7b` checkpoint directory (HF
    safetensors / .bin shards in the reference's key layout).  bitsandbytes paths are out of scope."""
    import os
    if load_8bit or load_4bit:
        raise NotImplementedError("load_8bit / load_4bit (bitsandbytes) are outside the hot-path scope (SURVEY.md §2)")
    if not os.path.isdir(str(model_path)):
        raise FileNotFoundError("load_pretrained_model: model_path must be a local checkpoint directory (no network here)")
    from transformers import AutoTokenizer
    from .vsm import config_from_hf, open_checkpoint
    tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)
    cfg = config_from_hf(model_path)
    engine = VQAEngine(VQAWeights(cfg, open_checkpoint(model_path, device=device), device=device))
    return tokenizer, LlavaSearchModel(engine, getattr(tokenizer, "eos_token_id", 2)), _Proc(), 2048


class VQA_LLM:
    """Same surface as /root/reference/vstar_bench_eval.py:38-165."""

    def __init__(self, args=None, engine: VQAEngine = None, tokenizer=None, conv_type="v1"):
        if engine is None:
            self.tokenizer, self.model, self.image_processor, self.context_len = load_pretrained_model(args.vqa_model_path, None, None)
            self.engine = self.model.engine
            self.conv_type = getattr(args, "conv_type", "v1")
            self.eos = getattr(self.tokenizer, "eos_token_id", 2)
            return
        self.engine = engine
        self.model = LlavaSearchModel(engine)
        self.tokenizer = tokenizer if tokenizer is not None else SyntheticTokenizer(engine.cfg)
        self.image_processor = _Proc()
        self.context_len = 2048
        self.conv_type = conv_type
        self.eos = getattr(self.tokenizer, "eos_token_id", 2)

    _img_src = None

    def get_patch(self, bbox, image_width, image_height, patch_size=224, patch_scale=None):
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

    def get_object_crop(self, image, bbox, patch_scale):
        b = self.get_patch(bbox, image.width, image.height, patch_scale=patch_scale)
        crop = image.crop((b[0], b[1], b[2], b[3])).resize((224, 224))
        return _clip_preprocess(crop)

    def _pixels(self, image, object_crops):
        if torch.is_tensor(image) and image.is_cuda:            # already preprocessed on the device (device_pixels)
            return image, object_crops
        img = _clip_preprocess(image).unsqueeze(0).to(self.engine.dev)
        img = ops.cast_f32_bf16(img.contiguous())
        crops = None
        if object_crops is not None and len(object_crops) > 0:
            crops = ops.cast_f32_bf16(torch.as_tensor(object_crops).float().to(self.engine.dev).contiguous())
        return img, crops

    def use_device_images(self, vsm):
        """share the VSM's resident search images and its Pillow-exact GPU resize pipeline: the padded image and the object
        crops of the option-scoring prompt (vstar_bench_eval.py:228-256) are then cut and resized on the device from the
        image that is already in HBM - bit-identical pixels, no host PIL work between the search and the answer"""
        self._img_src = vsm

    def device_pixels(self, image, boxes, patch_scale=1.2):
        """== (_clip_preprocess(expand2square_center(image)), [get_object_crop(image, b, patch_scale) for b in boxes]) as bf16
        device tensors ([1,3,224,224], [k,3,224,224] or None), computed by the GPU image pipeline"""
        vsm = self._img_src
        pipe = vsm.pipeline
        res = vsm.resident(image)                                   # uint8 [H, W, 3] in HBM
        H, W = int(res.shape[0]), int(res.shape[1])
        side = max(W, H)
        dev = self.engine.dev
        img = torch.empty((1, 3, 224, 224), dtype=BF, device=dev)
        if W == H:
            src = res
        else:                                                       # centred padding (vstar_bench_eval.py:25-36)
            bg = torch.tensor([int(x * 255) for x in self.image_processor.image_mean], dtype=torch.uint8, device=dev)
            src = bg.view(1, 1, 3).expand(side, side, 3).contiguous()
            left, top = (side - W) // 2, (side - H) // 2
            src[top:top + H, left:left + W] = res
        pipe._resize(src, 0, 0, side, side, side, side, 224, 224, out_bf16=img[0])
        crops = None
        if boxes is not None and len(boxes) > 0:
            crops = torch.empty((len(boxes), 3, 224, 224), dtype=BF, device=dev)
            for i, b in enumerate(boxes):
                l, t, r, btm = self.get_patch(b, W, H, patch_scale=patch_scale)
                pipe._resize(res, l, t, r - l, btm - t, r - l, btm - t, 224, 224, out_bf16=crops[i])
        return img, crops

    @torch.inference_mode()
    def free_form_inference(self, image, question, temperature=0, top_p=None, num_beams=1, max_new_tokens=200, object_crops=None,
                            images_long=None, objects_long=None):
        if temperature and temperature > 0:
            raise NotImplementedError("sampling is not on the V*Bench path (temperature=0, vstar_bench_eval.py:196)")
        prompt = build_prompt_v1(DEFAULT_IMAGE_TOKEN + "\n" + question)
        stop_str = "</s>"
        ids = tokenizer_image_object_token(prompt, self.tokenizer)
        kw = self.tokenizer(stop_str).input_ids
        if len(kw) > 1 and kw[0] == self.tokenizer.bos_token_id:
            kw = kw[1:]
        img, crops = self._pixels(image, object_crops)
        out = self.engine.generate(ids, img, crops, images_long, objects_long, max_new_tokens, self.eos, stop_ids=kw)
        text = self.tokenizer.batch_decode([out], skip_special_tokens=True)[0].strip()
        if text.endswith(stop_str):
            text = text[:-len(stop_str)]
        return text.strip()

    @torch.inference_mode()
    def free_form_inference_batch(self, images, questions, max_new_tokens=200, originals=None):
        """free_form_inference for several (image, question) pairs in ONE continuous-batched decode (not in the reference,
        which answers one image at a time, vstar_bench_eval.py:196): same prompt, stop string and post-processing per
        sample; each decode step reads the 7B weights once for the whole batch.  originals: the un-padded images the padded
        `images` were made from; with `use_device_images` the pixels are then produced on the GPU (bit-identical)."""
        stop_str = "</s>"
        kw = self.tokenizer(stop_str).input_ids
        if len(kw) > 1 and kw[0] == self.tokenizer.bos_token_id:
            kw = kw[1:]
        reqs = []
        for k, (image, question) in enumerate(zip(images, questions)):
            ids = tokenizer_image_object_token(build_prompt_v1(DEFAULT_IMAGE_TOKEN + "\n" + question), self.tokenizer)
            if originals is not None and self._img_src is not None:
                img, _ = self.device_pixels(originals[k], None)
            else:
                img, _ = self._pixels(image, None)
            reqs.append((ids, img, None, None, None))
        outs = self.engine.generate_batch(reqs, max_new_tokens, self.eos, stop_ids=kw)
        texts = []
        for out in outs:
            text = self.tokenizer.batch_decode([out], skip_special_tokens=True)[0].strip()
            if text.endswith(stop_str):
                text = text[:-len(stop_str)]
            texts.append(text.strip())
        return texts

    def _choice_item(self, image, question, options, object_crops, images_long, objects_long):
        qs = DEFAULT_IMAGE_TOKEN + "\n" + question
        q_ids = tokenizer_image_object_token(build_prompt_v1(qs), self.tokenizer)
        opt_ids = []
        for option in options:
            full = tokenizer_image_object_token(build_prompt_v1(qs, option), self.tokenizer)
            opt_ids.append(full[len(q_ids):])
        img, crops = self._pixels(image, object_crops)
        return (q_ids, opt_ids, img, crops, images_long, objects_long)

    @torch.inference_mode()
    def multiple_choices_inference(self, image, question, options, object_crops=None, images_long=None, objects_long=None):
        item = self._choice_item(image, question, options, object_crops, images_long, objects_long)
        return self.engine.option_losses_batch([item])[0][1]

    @torch.inference_mode()
    def multiple_choices_inference_batch(self, requests):
        """requests: [(image, question, options, object_crops, images_long, objects_long)] -> [chosen option index].
        Not in the reference (one sample at a time, vstar_bench_eval.py:257): same per-sample result, batched execution."""
        items = [self._choice_item(*r) for r in requests]
        return [c for _, c in self.engine.option_losses_batch(items)]
