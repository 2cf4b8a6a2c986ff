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
    # identical to oracle/make_golden.py:StubVSM (pure function of crop pixels).  hot: None = detection logits stay below
    # 0.45 (never confident); "root" = crops with min side >= 600 get three confident boxes (0.9 / 0.7 / 0.6); "small" = crops
    # with min side <= 300 get one confident box (0.8)
    def __init__(self, hot=None):
        self.calls = []
        self.hot = hot

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
        if self.hot == "root" and min(w, h) >= 600:
            logits[100, 0], logits[7, 0], logits[2000, 0] = 0.9, 0.7, 0.6
        if self.hot == "small" and min(w, h) <= 300:
            logits[55, 0] = 0.8
        return boxes, logits, hm


def synth_image(seed, w, h):
    from PIL import Image
    return Image.fromarray(np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")


# ---------------------------------------------------------------- V*Bench driver fixtures (tests/test_bench_eval.py)
MISSING_MSG = ("Sorry, I can not answer the question. Some visual information about the following objects is missing or "
               "unclear:")


class _Proc:
    image_mean = [0.48145466, 0.4578275, 0.40821073]


class StubVQA:
    """Deterministic stand-in for VQA_LLM (vstar_bench_eval.py:49-165): answers are pure functions of their arguments, so the
    option an implementation picks checks every string / box it fed in (focus message, normalised boxes, long/short flags)."""
    image_processor = _Proc()

    def __init__(self):
        self.log = []

    def free_form_inference(self, image, question, max_new_tokens=512):
        import zlib
        k = zlib.crc32(question.encode()) % 3
        self.log.append(("free", image.size, question))
        if k == 0:
            return "It is blue."
        if k == 1:
            return MISSING_MSG + " the red mug."
        return MISSING_MSG + " a dog, the blue umbrella, traffic light."

    def get_object_crop(self, image, bbox, patch_scale=1.0):
        return torch.tensor([round(float(v), 2) for v in bbox] + [patch_scale], dtype=torch.float32)

    def multiple_choices_inference(self, image, question, options, object_crops=None, images_long=None, objects_long=None):
        import zlib
        key = repr((image.size, question, options, None if object_crops is None else object_crops.round().tolist(),
                    images_long, objects_long))
        self.log.append(("choice", key))
        return zlib.crc32(key.encode()) % len(options)


def make_bench_folder(root):
    """tiny synthetic V*Bench tree: <root>/{direct_attributes,relative_position}/{name}.jpg|.png + {name}.json"""
    import json
    import os
    spec = {"direct_attributes": [("sa_1", 31, 640, 400), ("sa_2", 32, 300, 520), ("sa_3", 33, 512, 512), ("sa_4", 34, 700, 260)],
            "relative_position": [("sa_5", 35, 480, 360), ("sa_6", 36, 256, 600), ("sa_7", 37, 900, 300)]}
    for test_type, items in spec.items():
        d = os.path.join(root, test_type)
        os.makedirs(d, exist_ok=True)
        for name, seed, w, h in items:
            synth_image(seed, w, h).save(os.path.join(d, name + ".png"))
            with open(os.path.join(d, name + ".json"), "w") as f:
                json.dump({"question": f"What is the colour of item {seed}?", "options": ["red", "blue", "green", "white"][:2 + seed % 3]}, f)
    return root
