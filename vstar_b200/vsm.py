"""`VSM` — drop-in mirror of the reference's `visual_search.VSM` wrapper (/root/reference/visual_search.py:142-225)
on top of the sm_100a engine: host-side prompt building / tokenisation / image preprocessing, the batched engine call,
and the per-mode post-processing with the reference's return conventions.

Also hosts the checkpoint reader for the HF key layout and the synthetic tokenizer used when no
checkpoint/tokenizer exists (offline benches and tests).
"""
from __future__ import annotations

import glob
import json
import os
import zlib

import numpy as np
import torch

from . import ops
from .config import VSMConfig, IMAGE_TOKEN_INDEX
from .engine import VSMEngine, VSMWeights
from .image import GpuImagePipeline
from .visual_search import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_TOKEN, Heatmap, CudaScorer, _NodeEval)

BF = torch.bfloat16
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

LLAVA_V1_SYSTEM = ("A chat between a curious human and an artificial intelligence assistant. "
                   "The assistant gives helpful, detailed, and polite answers to the human's questions.")


def build_prompt(question, conv_type="llava_v1", use_mm_start_end=True):
    """visual_search.py:176-184 + conversation.py:355-365 (conv_llava_v1: SeparatorStyle.TWO, sep=' ', sep2='</s>')."""
    prompt = DEFAULT_IMAGE_TOKEN + "\n" + question
    if use_mm_start_end:
        prompt = prompt.replace(DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_TOKEN + DEFAULT_IM_END_TOKEN)
    if conv_type != "llava_v1":
        raise NotImplementedError("only the llava_v1 template is on the hot path (visual_search.py:48 default)")
    return LLAVA_V1_SYSTEM + " " + "USER" + ": " + prompt + " " + "ASSISTANT" + ":"


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX):
    """VisualSearch/model/llava/mm_utils.py:19-44, reused verbatim in behaviour (defines T and the 255 offset)."""
    prompt_chunks = [tokenizer(chunk).input_ids for chunk in prompt.split("<image>")]

    def insert_separator(X, sep):
        return [ele for sublist in zip(X, [sep] * len(X)) for ele in sublist][:-1]

    input_ids = []
    offset = 0
    if len(prompt_chunks) > 0 and len(prompt_chunks[0]) > 0 and prompt_chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        input_ids.append(prompt_chunks[0][0])
    for x in insert_separator(prompt_chunks, [image_token_index] * (offset + 1)):
        input_ids.extend(x[offset:])
    return input_ids


class SyntheticTokenizer:
    """Stand-in when no sentencepiece model exists offline: deterministic hash-word tokenizer with the special tokens
    of the VSM vocabulary (32000 Llama + [PAD] + [LOC] + <im_start> + <im_end>, VisualSearch/train.py:141-148)."""
    bos_token_id, eos_token_id, unk_token_id = 1, 2, 0

    def __init__(self, cfg: VSMConfig, pad_to=None):
        self.cfg = cfg
        self.loc = cfg.loc_token_idx
        self.im_start, self.im_end = cfg.vocab - 2, cfg.vocab - 1
        self.hi = min(cfg.vocab - 24, 31990)
        self.pad_to = pad_to

    def _word(self, w):
        return 3 + zlib.crc32(w.encode()) % (self.hi - 3)

    def __call__(self, text, add_special_tokens=True):
        ids = [self.bos_token_id] if add_special_tokens else []
        for tok in text.replace(DEFAULT_IM_START_TOKEN, " <im_start> ").replace(DEFAULT_IM_END_TOKEN, " <im_end> ").replace("[LOC]", " [LOC] ").split():
            ids.append({"<im_start>": self.im_start, "<im_end>": self.im_end, "[LOC]": self.loc}.get(tok, None) or self._word(tok))

        class _R:
            pass

        r = _R()
        r.input_ids = ids
        return r

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(f"w{int(t)}" for t in row if not (skip_special_tokens and int(t) in (0, 1, 2))) for row in ids]


# ----------------------------------------------------------------------------------------------------------------
# host image preprocessing (visual_search.py:186-194).  PIL bicubic like the HF processors the reference uses, into
# PINNED staging buffers so the H2D copy is asynchronous.
# ----------------------------------------------------------------------------------------------------------------
def expand2square(pil_img, background_color):
    # VisualSearch/utils/utils.py:28-39 : paste at TOP-LEFT, pad bottom/right
    from PIL import Image
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img, (0, 0))
    return result


_MEAN = np.array(CLIP_MEAN, np.float32)
_STD = np.array(CLIP_STD, np.float32)


def _normalize_into(arr_u8, dst):
    x = arr_u8.astype(np.float32) * np.float32(1 / 255.0)
    x = (x - _MEAN) / _STD
    dst.copy_(torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))))


def preprocess_clip_into(pil_img, dst, size=224):
    from PIL import Image
    bg = tuple(int(x * 255) for x in CLIP_MEAN)
    img = expand2square(pil_img.convert("RGB"), bg)
    w, h = img.size
    short = min(w, h)
    nw, nh = int(w * size / short), int(h * size / short)
    img = img.resize((nw, nh), resample=Image.BICUBIC)
    left, top = (nw - size) // 2, (nh - size) // 2
    _normalize_into(np.array(img.crop((left, top, left + size, top + size))), dst)


def preprocess_owl_into(pil_img, dst, size=768):
    from PIL import Image
    _normalize_into(np.array(pil_img.convert("RGB").resize((size, size), resample=Image.BICUBIC)), dst)


# ----------------------------------------------------------------------------------------------------------------
# checkpoint reader (HF sharded safetensors / .bin with the reference key layout)
# ----------------------------------------------------------------------------------------------------------------
def open_checkpoint(path):
    """-> callable name -> tensor, over *.safetensors or pytorch_model*.bin shards in `path`."""
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if files:
        from safetensors import safe_open
        index = {}
        handles = [safe_open(f, framework="pt", device="cpu") for f in files]
        for h in handles:
            for k in h.keys():
                index[k] = h
        return lambda name: index[name].get_tensor(name)
    files = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {path}")
    merged = {}
    for f in files:
        merged.update(torch.load(f, map_location="cpu", weights_only=True))
    return lambda name: merged[name]


def config_from_hf(path) -> VSMConfig:
    j = json.load(open(os.path.join(path, "config.json")))
    cfg = VSMConfig(hidden=j["hidden_size"], n_layers=j["num_hidden_layers"], n_heads=j["num_attention_heads"],
                    intermediate=j["intermediate_size"], vocab=j["vocab_size"], rms_eps=j.get("rms_norm_eps", 1e-6),
                    owl_query_dim=j.get("out_dim", 512))
    return cfg


class VSM:
    """Same surface as the reference wrapper: `VSM(args)`, `.inference(image, question, mode)`; plus `detect_batch`."""

    def __init__(self, args=None, engine: VSMEngine = None, tokenizer=None, frontier_batch=8, draft_answer="Sure, [LOC].",
                 forced_answer_ids=None, prep="gpu"):
        if engine is None:
            # visual_search.py:143-172: tokenizer + checkpoint from args.version, CLIP tower from args.vision_tower
            if args is None or not os.path.isdir(str(args.version)):
                raise FileNotFoundError(
                    "VSM(args): args.version must be a local checkpoint directory (no network here); pass engine= / tokenizer= "
                    "for synthetic weights")
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(args.version, model_max_length=args.model_max_length, padding_side="right", use_fast=False)
            tokenizer.pad_token = tokenizer.unk_token
            cfg = config_from_hf(args.version)
            cfg.loc_token_idx = tokenizer("[LOC]", add_special_tokens=False).input_ids[0]
            main = open_checkpoint(args.version)
            clip = open_checkpoint(args.vision_tower)        # CLIP weights are not in the VSM checkpoint (merge...py:146-149)

            def get(name):
                pfx = "model.vision_tower.vision_tower."
                return clip(name[len(pfx):]) if name.startswith(pfx) else main(name)

            engine = VSMEngine(VSMWeights(cfg, get))
            self.conv_type, self.use_mm_start_end = args.conv_type, args.use_mm_start_end
        else:
            self.conv_type, self.use_mm_start_end = "llava_v1", True
        self.engine = engine
        self.model = engine                       # attribute name kept from the reference wrapper
        self.cfg = engine.cfg
        self.vsm_tokenizer = tokenizer if tokenizer is not None else SyntheticTokenizer(self.cfg)
        self.frontier_batch = frontier_batch
        self.scorer = CudaScorer()
        eos = getattr(self.vsm_tokenizer, "eos_token_id", 2)
        self.eos = eos
        self.draft_ids = list(self.vsm_tokenizer(draft_answer, add_special_tokens=False).input_ids) + [eos]
        self.forced_answer_ids = forced_answer_ids      # synthetic weights only: logits-processor-style forcing
        if forced_answer_ids is not None:
            self.draft_ids = list(forced_answer_ids)
        self._pinned = {}
        self.timers = dict(prep=0.0, engine=0.0)
        # prep="gpu": crops are cut + resized on the device from the resident search image (Pillow-exact integer
        # bicubic); prep="host": the reference's PIL path (kept for parity tests of the pipeline itself)
        self.prep = prep
        self.pipeline = GpuImagePipeline(engine.dev, self.cfg.clip_image, self.cfg.owl_image) if prep == "gpu" else None
        self._resident = {}          # id(PIL image) -> (PIL image, uint8 HWC device tensor)
        # crops per engine call while search images are still being uploaded (see _run); 0 = never split
        self.upload_chunk = int(os.environ.get("VSB_UPLOAD_CHUNK", "16"))
        self.h2d_bytes = 0

    # ------------------------------------------------------------------ host prep
    def _staging(self, B):
        c = self.cfg
        if B not in self._pinned:
            self._pinned[B] = (torch.empty((B, 3, c.clip_image, c.clip_image), dtype=torch.float32).pin_memory(),
                               torch.empty((B, 3, c.owl_image, c.owl_image), dtype=torch.float32).pin_memory())
        return self._pinned[B]

    def _prep(self, images):
        """host path: PIL crops -> (clip [B,3,224,224], owl [B,3,768,768]) bf16 on the device"""
        c = self.cfg
        B = len(images)
        pc, po = self._staging(B)
        for i, im in enumerate(images):
            preprocess_clip_into(im, pc[i], c.clip_image)
            preprocess_owl_into(im, po[i], c.owl_image)
        self.h2d_bytes += pc.numel() * 4 + po.numel() * 4
        ic = ops.cast_f32_bf16(pc.cuda(non_blocking=True))      # .bfloat16() of the reference (visual_search.py:189,194)
        io = ops.cast_f32_bf16(po.cuda(non_blocking=True))
        return ic, io

    def resident(self, pil_img):
        """uint8 HWC copy of a search image in HBM (uploaded once, reused by every crop of that search)"""
        key = id(pil_img)
        hit = self._resident.get(key)
        if hit is None or hit[0] is not pil_img:
            if len(self._resident) >= 64:
                self._resident.pop(next(iter(self._resident)))
            t = self.pipeline.upload(pil_img)
            self.h2d_bytes += t.numel()
            self._resident[key] = (pil_img, t)
            return t
        return hit[1]

    def release(self, pil_img=None):
        if pil_img is None:
            self._resident.clear()
        else:
            self._resident.pop(id(pil_img), None)

    def _prep_regions(self, regions):
        """regions: list of (source PIL image, bbox [x,y,w,h]) -> device pixel tensors via the GPU image pipeline"""
        if self.prep != "gpu":
            crops = [src.crop((int(b[0]), int(b[1]), int(b[0] + b[2]), int(b[1] + b[3]))) for src, b in regions]
            return self._prep(crops)
        c = self.cfg
        B = len(regions)
        ic = torch.empty((B, 3, c.clip_image, c.clip_image), dtype=BF, device=self.engine.dev)
        io = torch.empty((B, 3, c.owl_image, c.owl_image), dtype=BF, device=self.engine.dev)
        for i, (src, b) in enumerate(regions):
            self.pipeline.crop_tensors(self.resident(src), b, ic[i], io[i])
        return ic, io

    def _ids(self, question):
        prompt = build_prompt(question, self.conv_type, self.use_mm_start_end)
        return tokenizer_image_token(prompt, self.vsm_tokenizer)

    # ------------------------------------------------------------------ engine calls
    def _run(self, regions, questions, mode):
        """regions: list of (source PIL image, bbox) -> list (per crop) of dicts with device tensors"""
        import time
        t0 = time.perf_counter()
        ids_list = [self._ids(q) for q in questions]
        groups = {}
        for i, ids in enumerate(ids_list):
            groups.setdefault((len(ids), ids.index(IMAGE_TOKEN_INDEX)), []).append(i)
        results = [None] * len(regions)
        pending = []
        for key, members in groups.items():
            # Search images that are not resident yet cost ~1.5 ms of host time each (PIL -> pinned staging -> H2D).  Such a
            # group is cut into chunks whose engine work is launched WITHOUT waiting for the GPU (defer=True), so the host
            # converts the next chunk's images while the GPU evaluates the previous one.
            fresh = sum(1 for i in members if self.prep == "gpu" and id(regions[i][0]) not in self._resident)
            step = self.upload_chunk if (0 < self.upload_chunk < min(fresh, len(members))) else len(members)
            for c0 in range(0, len(members), step):
                chunk = members[c0:c0 + step]
                ic, io = self._prep_regions([regions[i] for i in chunk])
                t1 = time.perf_counter()
                self.timers["prep"] += t1 - t0
                prompt = torch.tensor([ids_list[i] for i in chunk], dtype=torch.int64)
                out = self.engine.inference(io, ic, prompt, self.draft_ids, eos_token_id=self.eos, mode=mode,
                                            forced_ids=self.forced_answer_ids, defer=True)
                pending.append((chunk, out, io, ic, prompt))
                t0 = time.perf_counter()
                self.timers["engine"] += t0 - t1
        for chunk, out, io, ic, prompt in pending:
            self.engine.finish(out)
            bad = [j for j, ok in enumerate(out["verified"]) if not ok]
            for j, i in enumerate(chunk):
                if j in bad:
                    results[i] = self._exact_single(io[j:j + 1], ic[j:j + 1], prompt[j:j + 1], mode)
                else:
                    results[i] = self._slice(out, j, mode)
        self.timers["engine"] += time.perf_counter() - t0
        return results

    def _exact_single(self, io, ic, prompt, mode):
        """draft mismatch: exact greedy decode on the KV cache, then one teacher-forced pass over the emitted ids
        (== the reference's last generate step, VSM.py:459)."""
        out_ids, _ = self.engine.generate(prompt, ic, max_new_tokens=100, eos_token_id=self.eos)
        res = dict(output_ids=torch.tensor(out_ids))
        if mode == "vqa":
            return res
        ids = torch.tensor([out_ids[:-1]], dtype=torch.int64, device=self.engine.dev)
        if self.cfg.loc_token_idx not in out_ids:
            raise RuntimeError("no [LOC] token generated (reference: IndexError at visual_search.py:209-211)")
        out = self.engine.model_forward(io, ic, ids, mode=mode)
        out["output_ids"] = [res["output_ids"]]
        return self._slice(out, 0, mode)

    @staticmethod
    def _slice(out, j, mode):
        r = dict(output_ids=out["output_ids"][j])
        if mode == "vqa":
            return r
        locs = [k for k, b in enumerate(out["crop_of_loc"]) if b == j]
        r["low_res"] = out["low_res_masks"][locs[-1]]             # pred_mask[-1]  (visual_search.py:211, :225)
        if mode == "detection":
            f = locs[0]                                           # pred_boxes[0] / pred_logits[0]
            r["scores"] = out["scores"][f]
            r["logits"] = out["pred_logits"][f]
            r["boxes"] = out["pred_boxes"][f]                     # owl_heads returns boxes per (crop, query) entry
        return r

    # ------------------------------------------------------------------ public API
    @torch.inference_mode()
    def inference(self, image, question, mode="segmentation"):
        """visual_search.py:174-225.  'segmentation' -> Heatmap-backed fp32 tensor [h,w] on GPU (>= 0);
        'vqa' -> str; 'detection' -> (boxes [P,4] CPU, scores [P,1] CPU, heatmap [h,w] GPU)."""
        r = self._run([(image, [0, 0, image.width, image.height])], [question], mode)[0]
        if self.prep == "gpu":
            self.release(image)           # a crop handed in by a caller is not a long-lived search image
        if mode == "vqa":
            input_len = len(self._ids(question))
            text = self.vsm_tokenizer.batch_decode(r["output_ids"][input_len:].view(1, -1), skip_special_tokens=True)[0]
            return text.replace("\n", "").replace("  ", " ").strip()
        h, w = image.height, image.width
        heat = self.scorer.from_low_res(r["low_res"], h, w)
        if mode == "segmentation":
            return heat.map
        return r["boxes"].cpu(), r["scores"].view(-1, 1).cpu(), heat.map

    def detect_batch(self, images, questions):
        """PIL crops in, see detect_regions"""
        return self.detect_regions([(im, [0, 0, im.width, im.height]) for im in images], questions)

    @torch.inference_mode()
    def detect_regions(self, regions, questions):
        """Batched detection-mode evaluation for the search controller: regions = [(search image, bbox)] -> [_NodeEval]."""
        rs = self._run(regions, questions, "detection")
        sc = torch.stack([r["scores"] for r in rs])                          # [n, P]
        idx, val = ops.argmax_rows(sc.contiguous())
        bx = torch.stack([r["boxes"] for r in rs])                           # [n, P, 4]
        top = bx[torch.arange(len(rs), device=bx.device), idx.long()]
        host = torch.cat([val.view(-1, 1), top], dim=1).cpu()                # one D2H for the whole batch
        out = []
        for i, r in enumerate(rs):
            ev = _NodeEval()
            ev.n_logits = r["scores"].numel()
            ev.top_logit = float(host[i, 0])
            ev.top_box = host[i, 1:].clone()
            ev.boxes, ev.scores = r["boxes"], r["scores"].view(-1, 1)
            ev.low_res = r["low_res"]
            out.append(ev)
        return out


# ----------------------------------------------------------------------------------------------------------------
# model-loading API mirror: VSMForCausalLM.from_pretrained / .inference / .model_forward
# ----------------------------------------------------------------------------------------------------------------
class _VisionTowerHandle:
    """what `vsm_model.get_model().get_vision_tower()` hands to the reference wrapper (visual_search.py:160-162)"""

    def __init__(self):
        self.image_processor = None

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self


class _Cfg:
    pass


class VSMForCausalLM:
    """Drop-in for /root/reference/VisualSearch/model/VSM.py:162-553 (inference paths only) on the sm_100a engine."""

    def __init__(self, engine: VSMEngine, vision_tower_name="openai/clip-vit-large-patch14"):
        self.engine = engine
        self.cfg = engine.cfg
        self.loc_token_idx = engine.cfg.loc_token_idx
        self.config = _Cfg()
        self.config.vision_tower = vision_tower_name
        self.config.mm_vision_tower = vision_tower_name
        self.config.vocab_size = engine.cfg.vocab
        self._tower = _VisionTowerHandle()

    @classmethod
    def from_pretrained(cls, version, low_cpu_mem_usage=True, vision_tower="openai/clip-vit-large-patch14", loc_token_idx=None,
                        torch_dtype=torch.bfloat16, device_map="cuda", is_eval=True, **kwargs):
        """visual_search.py:157-159.  `version` = local checkpoint dir (HF shards, key layout of
        merge_lora_weights_and_save_hf_model.py:143-151); CLIP weights come from the `vision_tower` directory."""
        if torch_dtype != torch.bfloat16:
            raise NotImplementedError("the sm_100a kernels compute in bf16 (the reference loads the VSM in bf16 too)")
        cfg = config_from_hf(version)
        if loc_token_idx is not None:
            cfg.loc_token_idx = int(loc_token_idx)
        main = open_checkpoint(version)
        clip = open_checkpoint(vision_tower)

        def get(name):
            pfx = "model.vision_tower.vision_tower."
            return clip(name[len(pfx):]) if name.startswith(pfx) else main(name)

        device = "cuda" if device_map in ("cuda", "auto") else device_map
        return cls(VSMEngine(VSMWeights(cfg, get, device=device)), vision_tower)

    # -- HF-style accessors used by the reference wrapper
    def get_model(self):
        return self

    def initialize_vision_modules(self, cfg):
        return None                       # CLIP is part of the engine weights already

    def get_vision_tower(self):
        return self._tower

    def eval(self):
        return self

    def _ids_rows(self, ids_1d, T, img_pos):
        c = self.cfg
        return [(k - 1) + c.clip_tokens - 1 for k in (ids_1d == c.loc_token_idx).nonzero().flatten().tolist()]

    @torch.inference_mode()
    def model_forward(self, images, images_clip, input_ids, original_size_list=None, label_list=None, inference=True, **_unused):
        """VSM.py:201-364 with inference=True (batch of teacher-forced samples sharing one length)"""
        assert inference, "training losses are outside the hot path"
        out = self.engine.model_forward(images.to(torch.bfloat16), images_clip.to(torch.bfloat16), input_ids)
        sizes = original_size_list if original_size_list is not None else [tuple(l.shape) for l in label_list]
        pred_masks = []
        for b in range(input_ids.shape[0]):
            locs = [k for k, cb in enumerate(out["crop_of_loc"]) if cb == b]
            h, w = sizes[b]
            pred_masks.append(torch.stack([ops.heatmap(out["low_res_masks"][k].contiguous(), int(h), int(w), clamp=False, with_stats=False)[0] for k in locs]))
        return {"pred_masks": pred_masks, "gt_masks": None, "pred_logits": out["pred_logits"].unsqueeze(-1),
                "pred_boxes": out["pred_boxes"], "gt_bboxes": None}

    @torch.inference_mode()
    def inference(self, images_clip, images, input_ids, resize_list, original_size_list, max_new_tokens=32, tokenizer=None, mode="vqa"):
        """VSM.py:438-553: greedy generate, then (mode != 'vqa') the seg / det branches.  Returns
        (output_ids [1,L] | None, [pred_masks [n_loc,h,w]] | None, {'pred_logits','pred_boxes'} | None)."""
        assert mode in ["vqa", "segmentation", "detection"]
        assert input_ids.shape[0] == 1, "the reference wrapper evaluates one crop per call"
        eos = getattr(tokenizer, "eos_token_id", 2) if tokenizer is not None else 2
        ic = images_clip.to(torch.bfloat16)
        out_ids, _ = self.engine.generate(input_ids.cpu(), ic, max_new_tokens=max_new_tokens, eos_token_id=eos)
        output_ids = torch.tensor([out_ids], dtype=torch.int64, device=input_ids.device)
        if mode == "vqa":
            return output_ids, None, None
        if self.loc_token_idx not in out_ids:
            # the reference returns empty lists here and its wrapper then fails with IndexError (visual_search.py:209-211)
            return None, [], None
        ids = torch.tensor([out_ids[:-1]], dtype=torch.int64, device=self.engine.dev)     # last generate step's input (VSM.py:459)
        out = self.engine.model_forward(images.to(torch.bfloat16), ic, ids, mode=mode)
        h, w = original_size_list[0]
        pm = torch.stack([ops.heatmap(out["low_res_masks"][k].contiguous(), int(h), int(w), clamp=False, with_stats=False)[0]
                          for k in range(out["low_res_masks"].shape[0])])
        if mode == "segmentation":
            return None, [pm], None
        return None, [pm], {"pred_logits": out["pred_logits"].unsqueeze(-1), "pred_boxes": out["pred_boxes"]}
