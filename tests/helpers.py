"""Test-only helpers (NOT product code): a numpy heat-map scorer so the search controller's ordering logic can be
checked on a CPU-only box, and the deterministic stub VSM used by the committed search goldens."""
import numpy as np
import torch


class NumpyHeat:
    def __init__(self, arr):
        self.arr = arr                       # clamped [h,w] fp32
        self.h, self.w = arr.shape

    def host_stats(self):
        return np.array([self.arr.max(), self.arr.min(), self.arr.sum()], np.float32)

    def norm(self):
        if getattr(self, "_n", None) is None:
            mx, mn = self.arr.max(), self.arr.min()
            self._n = ((self.arr - mn) / (mx - mn) if mx != mn else self.arr * 0).astype(np.float64)
        return self._n

    def __array__(self, dtype=None, copy=None):
        return self.norm().reshape(self.h, self.w, 1)


class NumpyScorer:
    """restates ops.heatmap / ops.rect_sums with numpy (float64 sums of the normalised map)"""

    def from_low_res(self, low_res, h, w):
        t = torch.nn.functional.interpolate(low_res.float().cpu()[None, None], (h, w), mode="bilinear", align_corners=False)[0, 0]
        return NumpyHeat(t.clamp(min=0).numpy())

    def from_full_res(self, tensor, h, w):
        return NumpyHeat(tensor.reshape(h, w).float().cpu().clamp(min=0).numpy())

    def rect_sums(self, jobs):
        out = []
        for heat, rects in jobs:
            n = heat.norm()
            out.append(np.array([n[max(0, y):y + rh, max(0, x):x + rw].sum() for x, y, rw, rh in rects]))
        return out


class StubVSM:
    # identical to oracle/make_golden.py:StubVSM (pure function of crop pixels)
    def __init__(self):
        self.calls = []

    def inference(self, image, question, mode="segmentation"):
        arr = np.asarray(image, dtype=np.uint8)
        h, w = arr.shape[:2]
        self.calls.append((w, h, {"detection": 0, "vqa": 1, "segmentation": 2}[mode]))
        s = int(arr[::max(1, h // 16), ::max(1, w // 16)].astype(np.int64).sum()) % (2 ** 31)
        rng = np.random.default_rng(s)
        if mode == "vqa":
            return "The object is most likely to appear near the table."
        low = rng.standard_normal((12, 12)).astype(np.float32) * 4.0
        hm = torch.nn.functional.interpolate(torch.from_numpy(low)[None, None], (h, w), mode="bilinear",
                                             align_corners=False)[0, 0].clamp(min=0)
        if mode == "segmentation":
            return hm
        logits = torch.from_numpy(rng.uniform(0.0, 0.45, (2304, 1)).astype(np.float32))
        boxes = torch.from_numpy(rng.uniform(0.1, 0.9, (2304, 4)).astype(np.float32))
        return boxes, logits, hm


def synth_image(seed, w, h):
    from PIL import Image
    return Image.fromarray(np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
