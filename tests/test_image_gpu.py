"""GPU image pipeline (vsb_resample_h_u8 / vsb_resample_v_u8) against Pillow and the reference's host preprocessing:
uint8 results bit-exact with PIL.Image.resize(BICUBIC); bf16 pixel tensors bit-exact with
`preprocess(...).bfloat16()` of the PIL path (expand2square + resize + /255 + mean/std)."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def img(seed, w, h):
    return Image.fromarray(np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")


@pytest.mark.parametrize("w,h,ow,oh", [(150, 110, 224, 224), (233, 233, 224, 224), (1024, 1024, 224, 224), (700, 512, 768, 768),
                                       (96, 233, 768, 768), (2048, 1536, 768, 768), (64, 64, 64, 64)])
def test_resize_u8_bit_exact_vs_pil(w, h, ow, oh):
    from vstar_b200.image import GpuImagePipeline
    pipe = GpuImagePipeline()
    im = img(w + h, w, h)
    src = pipe.upload(im)
    out = torch.empty((oh, ow, 3), dtype=torch.uint8, device="cuda")
    pipe._resize(src, 0, 0, w, h, w, h, ow, oh, out_u8=out)
    ref = np.array(im.resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("W,H,bbox", [(640, 512, [0, 0, 640, 512]), (640, 512, [320, 256, 320, 256]), (1024, 1024, [512, 0, 512, 512]),
                                      (900, 1900, [0, 475, 900, 475]), (333, 517, [10, 20, 111, 400])])
def test_crop_tensors_equal_host_path(W, H, bbox):
    from oracle import vsm_oracle as O
    from vstar_b200.image import GpuImagePipeline
    pipe = GpuImagePipeline()
    im = img(W * 3 + H, W, H)
    src = pipe.upload(im)
    ic = torch.empty((3, 224, 224), dtype=BF, device="cuda")
    io = torch.empty((3, 768, 768), dtype=BF, device="cuda")
    pipe.crop_tensors(src, bbox, ic, io)
    crop = im.crop((bbox[0], bbox[1], bbox[0] + bbox[2], bbox[1] + bbox[3]))
    rc, ro = O.preprocess_clip(crop)[0].to(BF), O.preprocess_owl(crop)[0].to(BF)
    assert torch.equal(ic.cpu(), rc)
    assert torch.equal(io.cpu(), ro)


def test_vsm_gpu_prep_equals_host_prep():
    """the whole VSM call with prep='gpu' == prep='host' (same kernels downstream, identical pixel tensors)"""
    import json, os
    from oracle import vsm_oracle as O
    from vstar_b200.engine import VSMEngine, VSMWeights
    from vstar_b200.vsm import VSM
    j = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tiny_config.json")))
    cfg = O.VSMConfig(**j["cfg"])
    eng = VSMEngine(VSMWeights.from_state_dict(cfg, O.synthetic_state_dict(cfg, seed=j["weight_seed"])))
    prompt, ans = O.synthetic_prompt(cfg, n_text=24, seed=5)

    class V(VSM):
        def _ids(self, q):
            return prompt[0].tolist()

    a, b = V(engine=eng, forced_answer_ids=ans.tolist(), prep="gpu"), V(engine=eng, forced_answer_ids=ans.tolist(), prep="host")
    im = img(5, 300, 200)
    ba, sa, ha = a.inference(im, "q", mode="detection")
    bb, sb, hb = b.inference(im, "q", mode="detection")
    assert torch.equal(ba, bb) and torch.equal(sa, sb) and torch.equal(ha, hb)
