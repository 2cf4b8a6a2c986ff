"""One V*Bench sample through the whole SEAL loop — the caller of the hot path:
free-form answer -> "missing objects" -> guided visual search for each -> object crops -> option scoring.
Same per-sample logic and result keys as `eval_model` (/root/reference/vstar_bench_eval.py:186-273); the searches for all
missing objects of an image run in lock-step so their crop frontiers share GPU batches (the reference runs them one after
another, vstar_bench_eval.py:208-211).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch

from .visual_search import visual_search_many

MISSING_OBJECTS_MSG = ("Sorry, I can not answer the question. Some visual information about the following objects is missing or "
                       "unclear:")
FOCUS_MSG = "Additional visual information to focus on: "


def expand2square_center(pil_img, background_color):
    """vstar_bench_eval.py:25-36: CENTRED padding (the VSM wrapper pads bottom/right instead)"""
    from PIL import Image
    width, height = pil_img.size
    if width == height:
        return pil_img, 0, 0
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    left, top = (side - width) // 2, (side - height) // 2
    result.paste(pil_img, (left, top))
    return result, left, top


def normalize_bbox(bbox, image_width, image_height):
    nb = [bbox[0] / image_width, bbox[1] / image_height, (bbox[0] + bbox[2]) / image_width, (bbox[1] + bbox[3]) / image_height]
    return [np.clip(v, 0, 1) for v in nb]


def parse_missing_objects(prediction):
    if MISSING_OBJECTS_MSG not in prediction:
        return []
    tail = prediction.split(MISSING_OBJECTS_MSG)[-1]
    if tail.endswith("."):
        tail = tail[:-1]
    return [t.strip() for t in tail.split(",")]


def focus_question(question, object_names, norm_boxes):
    msg = FOCUS_MSG
    for i, (name, b) in enumerate(zip(object_names, norm_boxes)):
        msg += "{} <object> at location [{:.3f},{:.3f},{:.3f},{:.3f}]".format(name, b[0], b[1], b[2], b[3])
        msg += "; " if i != len(norm_boxes) - 1 else "."
    return msg + "\n" + question


def smallest_size_for(image, minimum_size_scale=4.0, minimum_size=224):
    """vstar_bench_eval.py:210"""
    return max(int(np.ceil(min(image.width, image.height) / minimum_size_scale)), minimum_size)


def collect_search_results(missing, results):
    """vstar_bench_eval.py:212-225: boxes of the final patch shifted to image coordinates, one entry per found instance"""
    search_result = []
    for name, (final_step, _path_length, _ok, all_valid) in zip(missing, results):
        patch = final_step["bbox"]
        boxes = all_valid if all_valid is not None else [final_step["detection_result"]]
        for b in boxes:
            b = b.clone()
            b[0] += patch[0]
            b[1] += patch[1]
            search_result.append({"bbox": b.tolist(), "name": name})
    return search_result


def choose_option(vqa_llm, image, question, options, missing, search_result):
    """vstar_bench_eval.py:226-257: option scoring, with the searched objects spliced in as <object> features"""
    return vqa_llm.multiple_choices_inference(*option_request(vqa_llm, image, question, options, missing, search_result))


def choose_options(vqa_llm, samples):
    """choose_option for several samples [(image, question, options, missing, search_result)]: one batched option scoring when
    the VQA LLM offers it (`multiple_choices_inference_batch`), the reference's one-at-a-time calls otherwise"""
    reqs = [option_request(vqa_llm, *s) for s in samples]
    if hasattr(vqa_llm, "multiple_choices_inference_batch") and len(reqs) > 1:
        return vqa_llm.multiple_choices_inference_batch(reqs)
    return [vqa_llm.multiple_choices_inference(*r) for r in reqs]


def option_request(vqa_llm, image, question, options, missing, search_result):
    """-> the argument tuple (image, question, options, object_crops, images_long, objects_long) of multiple_choices_inference"""
    if not missing:
        # the reference scores the options on the re-opened, UNPADDED image here (vstar_bench_eval.py:227, :257)
        return (image, question, options, None, None, None)
    names = [r["name"] for r in search_result]
    boxes = deepcopy([r["bbox"] for r in search_result])
    objects_long = [True] * len(names) if len(names) <= 2 else [False] * len(names)
    side = max(image.width, image.height)
    left, top = (side - image.width) // 2, (side - image.height) // 2          # expand2square (vstar_bench_eval.py:25-36)
    shifted = [[b[0] + left, b[1] + top, b[2], b[3]] for b in boxes]
    nboxes = [normalize_bbox(b, side, side) for b in shifted]
    if getattr(vqa_llm, "_img_src", None) is not None:
        # padded image + object crops cut and resized on the GPU from the resident search image (bit-identical pixels)
        padded, crops = vqa_llm.device_pixels(image, boxes, patch_scale=1.2)
    else:
        bg = tuple(int(x * 255) for x in vqa_llm.image_processor.image_mean)
        padded, _, _ = expand2square_center(image, bg)
        crops = torch.stack([vqa_llm.get_object_crop(image, b, patch_scale=1.2) for b in boxes], 0)
    return (padded, focus_question(question, names, nboxes), options, crops, [False], objects_long)


def seal_answer(vqa_llm, vsm, image, question, options, minimum_size_scale=4.0, minimum_size=224, search_batch=16,
                prediction_override=None, search_kwargs=None):
    """-> dict with the reference's per-sample result keys (question, options, prediction_freeform, missing_objects,
    search_result, option_chosen)"""
    bg = tuple(int(x * 255) for x in vqa_llm.image_processor.image_mean)
    padded, _, _ = expand2square_center(image, bg)
    prediction = prediction_override if prediction_override is not None else vqa_llm.free_form_inference(padded, question)
    missing = parse_missing_objects(prediction)
    search_result = []
    if missing:
        smallest = smallest_size_for(image, minimum_size_scale, minimum_size)
        jobs = [(image, name, smallest) for name in missing]
        results, _ = visual_search_many(vsm, jobs, batch_size=search_batch, **(search_kwargs or {}))
        search_result = collect_search_results(missing, results)
    chosen = choose_option(vqa_llm, image, question, options, missing, search_result)
    return dict(question=question, options=options, prediction_freeform=prediction, missing_objects=missing,
                search_result=search_result, option_chosen=chosen, correct=1 if chosen == 0 else 0)
