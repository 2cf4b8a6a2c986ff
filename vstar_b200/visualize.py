"""Search-path visualisation — mirror of /root/reference/visual_search.py:285-376 (`visualize_bbox`, `show_heatmap_on_image`,
`vis_heatmap`, `visualize_search_path`): the same files in `save_path` (whole_image.jpg, step_{i}.jpg, step_{i}_heatmap.jpg,
final_patch_image.jpg, search_result.jpg, context_cue.txt) drawn with the same OpenCV calls, from the reference-compatible
`search_path` fields (`bbox`, `detection_result`, `final_heatmap`, `context_cue`).

Not on the hot path: `final_heatmap` of a node is materialised here (np.asarray of the lazy heat-map object, visual_search.py)
only for the steps that are drawn.  The reference imports matplotlib next to cv2 but never uses it for these files.
"""
from __future__ import annotations

import os

import numpy as np

BOX_COLOR = (255, 0, 0)       # red (RGB order until the final cvtColor)
TEXT_COLOR = (255, 255, 255)


def visualize_bbox(img, bbox, class_name, color=BOX_COLOR, thickness=2):
    """rectangle + filled label strip above its top-left corner (visual_search.py:289-306)"""
    import cv2
    x_min, y_min, w, h = bbox
    x_min, x_max, y_min, y_max = int(x_min), int(x_min + w), int(y_min), int(y_min + h)
    cv2.rectangle(img, (x_min, y_min), (x_max, y_max), color=color, thickness=thickness)
    (tw, th), _ = cv2.getTextSize(class_name, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 1)
    cv2.rectangle(img, (x_min, y_min - int(1.3 * th)), (x_min + tw, y_min), BOX_COLOR, -1)
    cv2.putText(img, text=class_name, org=(x_min, y_min - int(0.3 * th)), fontFace=cv2.FONT_HERSHEY_SIMPLEX, fontScale=0.5,
                color=TEXT_COLOR, lineType=cv2.LINE_AA)
    return img


def show_heatmap_on_image(img, mask, use_rgb=False, colormap=None, image_weight=0.5):
    """JET overlay of a [0,1] mask on a [0,1] float image, renormalised to its maximum (visual_search.py:307-330)"""
    import cv2
    mask = np.clip(mask, 0, 1)
    heatmap = cv2.applyColorMap(np.uint8(255 * mask), cv2.COLORMAP_JET if colormap is None else colormap)
    if use_rgb:
        heatmap = cv2.cvtColor(heatmap, cv2.COLOR_BGR2RGB)
    heatmap = np.float32(heatmap) / 255
    if np.max(img) > 1:
        raise Exception("The input image should np.float32 in the range [0, 1]")
    if image_weight < 0 or image_weight > 1:
        raise Exception(f"image_weight should be in the range [0, 1]. Got: {image_weight}")
    cam = (1 - image_weight) * heatmap + image_weight * img
    cam = cam / np.max(cam)
    return np.uint8(255 * cam)


def vis_heatmap(image, heatmap, use_rgb=False):
    hi, lo = np.max(heatmap), np.min(heatmap)
    if hi != lo:
        heatmap = (heatmap - lo) / (hi - lo)
    return show_heatmap_on_image(image.astype(float) / 255., heatmap, use_rgb=use_rgb)


def _crop(image, box):
    return image.crop((box[0], box[1], box[0] + box[2], box[1] + box[3]))


def visualize_search_path(image, search_path, search_length, target_bbox, label, save_path):
    """visual_search.py:339-376"""
    import cv2
    os.makedirs(save_path, exist_ok=True)
    image.save(os.path.join(save_path, "whole_image.jpg"))
    whole = np.array(image)
    if target_bbox is not None:
        whole = visualize_bbox(whole.copy(), target_bbox, class_name="gt: " + label, color=(255, 0, 0))
    cues = []
    for step_i, node in enumerate(search_path):
        if step_i + 1 > search_length:
            break
        box = node["bbox"]
        if "detection_result" in node:
            final_patch = _crop(image, box)
            final_patch.save(os.path.join(save_path, "final_patch_image.jpg"))
            drawn = visualize_bbox(np.array(final_patch), [float(v) for v in node["detection_result"]], class_name="search result",
                                   color=(255, 0, 0))
            cv2.imwrite(os.path.join(save_path, "search_result.jpg"), cv2.cvtColor(drawn, cv2.COLOR_RGB2BGR))
        cur = visualize_bbox(whole.copy(), box, class_name="step-{}".format(step_i + 1), color=(0, 0, 255))
        cv2.imwrite(os.path.join(save_path, "step_{}.jpg".format(step_i + 1)), cv2.cvtColor(cur, cv2.COLOR_RGB2BGR))
        if "context_cue" in node:
            cues.append("step{}: {}".format(step_i + 1, node["context_cue"]) + "\n")
        if "final_heatmap" in node:
            score_map = np.asarray(node["final_heatmap"])          # lazy heat-map -> the reference's [h,w,1] fp32 array
            overlay = vis_heatmap(np.array(_crop(image, box)), score_map, use_rgb=True)
            cv2.imwrite(os.path.join(save_path, "step_{}_heatmap.jpg".format(step_i + 1)), cv2.cvtColor(overlay, cv2.COLOR_RGB2BGR))
    with open(os.path.join(save_path, "context_cue.txt"), "w") as f:
        f.writelines(cues)
