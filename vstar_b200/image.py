"""GPU image pipeline for the crop frontier: the search image stays resident in HBM as uint8 HWC and
every crop's CLIP (224x224, expand2square + bicubic) and OWL-ViT (768x768, bicubic) pixel tensors are produced on the
device by `vsb_resample_h_u8` / `vsb_resample_v_u8` — bit-identical to the Pillow resize the reference's HF processors
call (/root/reference/visual_search.py:186-194), without the per-crop PIL crop/deepcopy/resize and the 7 MB fp32 H2D.

The coefficient tables are computed here on the host exactly like Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc`
(libImaging/Resample.c): double arithmetic, bicubic a = -0.5, support = 2 * max(1, in/out), 22-bit fixed point.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import call

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_MEAN_C = (ctypes.c_float * 3)(*np.array(CLIP_MEAN, np.float32))
_STD_C = (ctypes.c_float * 3)(*np.array(CLIP_STD, np.float32))
BG = tuple(int(x * 255) for x in CLIP_MEAN)          # expand2square background (visual_search.py:186)

_coef_cache = {}


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    r = np.zeros_like(x)
    m1 = x < 1.0
    m2 = (x >= 1.0) & (x < 2.0)
    r[m1] = ((a + 2.0) * x[m1] - (a + 3.0)) * x[m1] * x[m1] + 1
    r[m2] = (((x[m2] - 5) * x[m2] + 8) * x[m2] - 4) * a
    return r


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """-> (coefs int32 [out, ksize], bounds int32 [out, 2] = (xmin, count), ksize) for a full-extent resize in_size -> out_size"""
    key = (in_size, out_size)
    if key in _coef_cache:
        return _coef_cache[key]
    scale = filterscale = float(in_size) / float(out_size)
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    coefs = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        w = _bicubic((x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                      # sequential double accumulation like the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        q = np.where(w < 0, -0.5 + w * (1 << 22), 0.5 + w * (1 << 22))
        coefs[xx, :xmax] = np.trunc(q).astype(np.int64).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    _coef_cache[key] = (coefs, bounds, ksize)
    return _coef_cache[key]


class _DevCoefs:
    """device copies of the coefficient tables, cached per (in, out)"""

    def __init__(self, device):
        self.device = device
        self.cache = {}

    def get(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self.cache:
            c, b, k = pil_bicubic_coeffs(in_size, out_size)
            self.cache[key] = (torch.from_numpy(c).to(self.device), torch.from_numpy(b).to(self.device), k)
            if len(self.cache) > 4096:
                self.cache.pop(next(iter(self.cache)))
        return self.cache[key]


class GpuImagePipeline:
    def __init__(self, device="cuda", clip_size=224, owl_size=768):
        self.device = device
        self.clip_size, self.owl_size = clip_size, owl_size
        self.coefs = _DevCoefs(device)
        self._tmp = None
        self._ring = []              # pinned staging slots: [tensor, numpy view, event of the last copy out of it]
        self._ring_next = 0
        self._copy_stream = None

    RING_SLOTS = 8

    def upload(self, pil_img):
        """PIL RGB image -> resident uint8 [H, W, 3] device tensor (one H2D per search image).

        The copy runs on a side stream out of a small ring of reusable pinned staging buffers (page-locking a fresh buffer
        per image costs more than the copy itself) and the compute stream waits for it with an event, never the host: the
        host only blocks when the ring wraps onto a slot whose copy has not finished.  That lets the caller convert the
        next image while earlier crops are being evaluated (VSM._run pipelines chunks this way)."""
        if pil_img.mode != "RGB":
            pil_img = pil_img.convert("RGB")
        w, h = pil_img.size
        n = h * w * 3
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        limit = self.RING_SLOTS if n <= (32 << 20) else 2     # huge search images (8192^2 = 201 MB): two staging buffers only
        if len(self._ring) < limit:
            self._ring.append([None, None, None])
        k = self._ring_next % min(len(self._ring), limit)
        self._ring_next += 1
        slot = self._ring[k]
        if slot[2] is not None:
            slot[2].synchronize()                 # the previous copy out of this slot has finished
        if slot[0] is None or slot[0].numel() < n:
            slot[0] = torch.empty(max(n, 4 << 20), dtype=torch.uint8).pin_memory()
            slot[1] = slot[0].numpy()
        np.copyto(slot[1][:n].reshape(h, w, 3), np.asarray(pil_img))
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            dev = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)     # from the copy stream's pool
            dev.copy_(slot[0][:n].view(h, w, 3), non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(self._copy_stream)
        dev.record_stream(main)                   # consumed by kernels on the compute stream
        main.wait_event(evt)
        slot[2] = evt
        return dev

    def _scratch(self, nbytes):
        if self._tmp is None or self._tmp.numel() < nbytes:
            self._tmp = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._tmp

    def _resize(self, src, x0, y0, cw, ch, in_w, in_h, out_w, out_h, out_bf16=None, out_u8=None, out_f32=None):
        """virtual input = crop (cw x ch at x0,y0) padded bottom/right with BG to in_w x in_h; Pillow resize to out_w x out_h"""
        st = torch.cuda.current_stream().cuda_stream
        ch_c, ch_b, ch_k = self.coefs.get(in_w, out_w)
        cv_c, cv_b, cv_k = self.coefs.get(in_h, out_h)
        tmp = self._scratch(in_h * out_w * 3)
        _lib.launches += 2
        call("vsb_resample_h_u8", src.data_ptr(), src.stride(0), x0, y0, cw, ch, in_h, BG[0], BG[1], BG[2], ch_c.data_ptr(), ch_b.data_ptr(),
             ch_k, out_w, tmp.data_ptr(), st)
        call("vsb_resample_v_u8", tmp.data_ptr(), out_w, cv_c.data_ptr(), cv_b.data_ptr(), cv_k, out_h,
             0 if out_u8 is None else out_u8.data_ptr(), 0 if out_bf16 is None else out_bf16.data_ptr(),
             0 if out_f32 is None else out_f32.data_ptr(), _MEAN_C, _STD_C, st)

    def crop_tensors(self, src, bbox, out_clip, out_owl):
        """src uint8 [H,W,3] on device; bbox = [x, y, w, h] (ints, as image.crop((x, y, x+w, y+h)) in visual_search.py:394);
        writes bf16 CHW pixel tensors into out_clip [3,224,224] and out_owl [3,768,768]."""
        x0, y0 = int(bbox[0]), int(bbox[1])
        cw, ch = int(bbox[0] + bbox[2]) - x0, int(bbox[1] + bbox[3]) - y0
        side = max(cw, ch)
        # CLIP: expand2square (pad bottom/right) -> shortest edge 224 bicubic (square => 224x224) -> centre crop (no-op)
        self._resize(src, x0, y0, cw, ch, side, side, self.clip_size, self.clip_size, out_bf16=out_clip)
        # OWL-ViT: the un-padded crop stretched to 768x768
        self._resize(src, x0, y0, cw, ch, cw, ch, self.owl_size, self.owl_size, out_bf16=out_owl)
