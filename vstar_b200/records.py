"""Crop records: the fixed-size result of ONE crop evaluation as the search controller consumes it.

A record is what crosses PCIe (one D2H per frontier batch) and NVLink (one all-gather per batch when the frontier is sharded
over GPUs): a few hundred floats instead of the reference's 2304 x 5 detections + the H x W heat-map per node
(/root/reference/visual_search.py:208-225, :448).  Layout = csrc/heads.cu (`REC_*`), include/vstar_b200.h:

    [0]      best sigmoid score                      (pred_logits.view(-1).max(),       visual_search.py:400)
    [1..4]   its box, cxcywh in (0,1)                (pred_bboxes[top_index],           :401)
    [5]      number of rows P (len(pred_logits))     (:398)
    [6]      rows with score > 0.5                   (all_valid_boxes mask,             :407)
    [7]      best row index, -1 if no finite score
    [8..10]  (max, min, sum) of the clamped H x W target-cue map (score_max,            :420)
    [11]     number of rectangle sums that follow
    [12..75] the first 16 boxes with score > 0.5, row order
    [76..]   sums of normalize_score(map) over the crop (first) and every quad-tree descendant of the crop, in
             `pyramid_rects` order — the per-ancestor terms of get_subpatch_scores (:255-266, :453-462)

Host-side pieces only (INT geometry + parsing); the numbers are produced by `ops.pack_detections` / `ops.heat_pyramids`.
"""
from __future__ import annotations

import numpy as np

REC_TOP, REC_BOX, REC_NROWS, REC_NVALID, REC_TOPIDX = 0, 1, 5, 6, 7
REC_MAX, REC_MIN, REC_SUM, REC_NRECT, REC_VALID, REC_MAXVALID, REC_PYR = 8, 9, 10, 11, 12, 16, 76


def record_floats(n_rects: int) -> int:
    """record length in floats for up to n_rects rectangle sums (multiple of 4 floats = 16 bytes)"""
    return (REC_PYR + max(0, int(n_rects)) + 3) // 4 * 4


def split_4subpatches(bbox):
    # /root/reference/visual_search.py:234-241
    r = bbox[3] / bbox[2]
    return (1, 4) if r >= 2 else (4, 1) if r <= 0.5 else (2, 2)


def get_sub_patches(bbox, nw, nh):
    # /root/reference/visual_search.py:243-253 (INT geometry: the last child takes the remainder)
    ws, hs = int(bbox[2] // nw), int(bbox[3] / nh)
    out = []
    for j in range(nh):
        for i in range(nw):
            out.append([bbox[0] + i * ws, bbox[1] + j * hs,
                        bbox[2] - i * ws if i == nw - 1 else ws, bbox[3] - j * hs if j == nh - 1 else hs])
    return out, ws, hs


def expandable(bbox, smallest_size) -> bool:
    """a node is split iff it is larger than the smallest unit (visual_search.py:416-417)"""
    return not (min(bbox[2], bbox[3]) <= smallest_size)


_PYR_CACHE = {}


def pyramid_rects(bbox, smallest_size):
    """Absolute rectangles whose sums a node's heat-map contributes to the search: the node itself (index 0), then, breadth
    first, the children of every expandable node of the quad-tree below it.  Empty for a node that is never split."""
    key = (tuple(int(v) for v in bbox), smallest_size)
    hit = _PYR_CACHE.get(key)
    if hit is not None:
        return hit
    out = []
    if expandable(bbox, smallest_size):
        out.append(key[0])
        level = [list(key[0])]
        while level:
            nxt = []
            for b in level:
                subs, _, _ = get_sub_patches(b, *split_4subpatches(b))
                for s in subs:
                    out.append(tuple(int(v) for v in s))
                    if expandable(s, smallest_size):
                        nxt.append(s)
            level = nxt
    if len(_PYR_CACHE) > 4096:
        _PYR_CACHE.clear()
    _PYR_CACHE[key] = out
    return out


class RecordPyramid:
    """rectangle sums of one node's normalised map, looked up by absolute rectangle"""
    __slots__ = ("stats", "sums")

    def __init__(self, stats, rects, values):
        self.stats = stats                                  # numpy float32 [3]: max, min, sum
        self.sums = dict(zip(rects, values))

    def missing(self, rects):
        return [r for r in rects if r not in self.sums]

    def get(self, rect):
        return self.sums[rect]


def parse_record(row, bbox, smallest_size):
    """row: numpy float32 [R] -> dict of host values (see module docstring)"""
    n_valid = int(row[REC_NVALID])
    n_rect = int(row[REC_NRECT])
    top_idx = int(row[REC_TOPIDX])
    out = dict(top_logit=float(row[REC_TOP]), top_box=row[REC_BOX:REC_BOX + 4].copy(), n_logits=int(row[REC_NROWS]), n_valid=n_valid,
               top_index=top_idx, valid_boxes=row[REC_VALID:REC_VALID + 4 * min(n_valid, REC_MAXVALID)].reshape(-1, 4).copy(),
               pyramid=None)
    if n_rect:
        rects = pyramid_rects(bbox, smallest_size)
        assert len(rects) == n_rect, (len(rects), n_rect, bbox, smallest_size)
        out["pyramid"] = RecordPyramid(row[REC_MAX:REC_MAX + 3].copy(), rects, row[REC_PYR:REC_PYR + n_rect].astype(np.float64))
    return out
