"""Host half of the GPU image pipeline: the Pillow-compatible bicubic coefficient tables (vstar_b200/image.py) drive an
integer two-pass resample that must equal PIL.Image.resize(BICUBIC) bit for bit (numpy emulation of the kernels here;
the CUDA kernels themselves are checked against PIL in tests/test_image_gpu.py).  Also pins the oracle's PIL-based
preprocessing against the HF PIL image processors."""
import numpy as np
import pytest
import torch
from PIL import Image

from vstar_b200.image import pil_bicubic_coeffs


def emulate(arr, out_w, out_h):
    h, w, _ = arr.shape
    ch, bh, kh = pil_bicubic_coeffs(w, out_w)
    cv, bv, kv = pil_bicubic_coeffs(h, out_h)
    tmp = np.zeros((h, out_w, 3), np.uint8)
    a = arr.astype(np.int64)
    for xo in range(out_w):
        x0, n = bh[xo]
        acc = (1 << 21) + (a[:, x0:x0 + n, :] * ch[xo, :n].astype(np.int64)[None, :, None]).sum(1)
        tmp[:, xo, :] = np.clip(acc >> 22, 0, 255)
    out = np.zeros((out_h, out_w, 3), np.uint8)
    t = tmp.astype(np.int64)
    for yo in range(out_h):
        y0, n = bv[yo]
        acc = (1 << 21) + (t[y0:y0 + n] * cv[yo, :n].astype(np.int64)[:, None, None]).sum(0)
        out[yo] = np.clip(acc >> 22, 0, 255)
    return out


@pytest.mark.parametrize("w,h,ow,oh", [(150, 110, 224, 224), (233, 233, 224, 224), (1024, 1024, 224, 224), (700, 512, 768, 768),
                                       (96, 233, 768, 768), (1500, 1500, 224, 224), (64, 64, 64, 64)])
def test_coeffs_reproduce_pil_bicubic(w, h, ow, oh):
    arr = np.random.default_rng(w * 7 + h).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(arr, "RGB").resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(emulate(arr, ow, oh), ref)


def test_oracle_preprocess_equals_hf_pil_processors():
    """oracle.preprocess_* (== vstar_b200.vsm host path) vs transformers' PIL-backed processors (the reference pins
    transformers 4.31 whose processors are PIL-based): fp32 rounding only."""
    try:
        from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
        from transformers.models.owlvit.image_processing_pil_owlvit import OwlViTImageProcessorPil
    except Exception:
        pytest.skip("PIL-backed HF processors not available")
    from oracle import vsm_oracle as O
    for (w, h) in [(150, 110), (96, 233), (512, 512)]:
        img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
        cp = CLIPImageProcessorPil()
        bg = tuple(int(x * 255) for x in cp.image_mean)
        a = cp.preprocess(O.expand2square(img, bg), return_tensors="pt")["pixel_values"]
        assert float((a - O.preprocess_clip(img)).abs().max()) < 2e-6
        b = OwlViTImageProcessorPil()(images=np.array(img), return_tensors="pt")["pixel_values"]
        assert float((b - O.preprocess_owl(img)).abs().max()) < 2e-6
