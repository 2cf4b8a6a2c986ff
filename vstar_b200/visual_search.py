"""Drop-in mirror of /root/reference/visual_search.py for the guided visual-search hot path.

Same public names and conventions: `parse_args`, `VSM`, `VSM.inference(image, question, mode)`,
`visual_search(...) -> (final_step, path_length, search_successful, all_valid_boxes)`, plus the INT geometry helpers.

What is different underneath (B200-first):
  * the controller is ITERATIVE (the reference recurses once per expanded node and dies at ~990 nodes)
    but pops the very same `queue.PriorityQueue` / `Prioritize` objects in the very same push/pop sequence, so the
    expansion order (including heapq tie behaviour) is the reference's;
  * frontier nodes are evaluated SPECULATIVELY in batches (each node's outputs are a pure function of its crop) and
    committed strictly in pop order; speculative work is simply dropped when the search ends early;
  * several searches can run in lock-step (`visual_search_many`) so their frontiers share one GPU batch;
  * heatmap statistics and the ancestor-chain sub-patch sums run on the GPU (ops.heatmap / ops.rect_sums); only a
    handful of scalars cross PCIe per expansion instead of the H x W fp32 map (visual_search.py:448).
"""
from __future__ import annotations

import argparse
import functools
from queue import PriorityQueue

import numpy as np
import torch

from . import ops

# ---------------------------------------------------------------------------------------------------------------
# CLI / constants (visual_search.py:28-52; VisualSearch/utils/utils.py:7-12)
# ---------------------------------------------------------------------------------------------------------------
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"

DETECTION_QUESTION = "Please locate the {} in this image."
CUE_QUESTION = ("According to the common sense knowledge and possible visual cues, what is the most likely location of "
                "the {} in the image?")


def parse_args(args):
    parser = argparse.ArgumentParser(description="Visual Search Evaluation")
    parser.add_argument("--version", default="craigwu/seal_vsm_7b")
    parser.add_argument("--benchmark-folder", default="vstar_bench", type=str)
    parser.add_argument("--visualization", action="store_true", default=False)
    parser.add_argument("--output_path", default="", type=str)
    parser.add_argument("--confidence_low", default=0.3, type=float)
    parser.add_argument("--confidence_high", default=0.5, type=float)
    parser.add_argument("--target_cue_threshold", default=6.0, type=float)
    parser.add_argument("--target_cue_threshold_decay", default=0.7, type=float)
    parser.add_argument("--target_cue_threshold_minimum", default=3.0, type=float)
    parser.add_argument("--minimum_size_scale", default=4.0, type=float)
    parser.add_argument("--minimum_size", default=224, type=int)
    parser.add_argument("--model_max_length", default=512, type=int)
    parser.add_argument("--vision-tower", default="openai/clip-vit-large-patch14", type=str)
    parser.add_argument("--use_mm_start_end", action="store_true", default=True)
    parser.add_argument("--conv_type", default="llava_v1", type=str, choices=["llava_v1", "llava_llama_2"])
    return parser.parse_args(args)


# ---------------------------------------------------------------------------------------------------------------
# INT geometry (visual_search.py:227-253, :277-283) — host code, bit-exact by construction
# ---------------------------------------------------------------------------------------------------------------
def refine_bbox(bbox, image_width, image_height):
    bbox[0] = max(0, bbox[0])
    bbox[1] = max(0, bbox[1])
    bbox[2] = min(bbox[2], image_width - bbox[0])
    bbox[3] = min(bbox[3], image_height - bbox[1])
    return bbox


def split_4subpatches(current_patch_bbox):
    hw_ratio = current_patch_bbox[3] / current_patch_bbox[2]
    if hw_ratio >= 2:
        return 1, 4
    elif hw_ratio <= 0.5:
        return 4, 1
    else:
        return 2, 2


def get_sub_patches(current_patch_bbox, num_of_width_patches, num_of_height_patches):
    width_stride = int(current_patch_bbox[2] // num_of_width_patches)
    height_stride = int(current_patch_bbox[3] / num_of_height_patches)
    sub_patches = []
    for j in range(num_of_height_patches):
        for i in range(num_of_width_patches):
            sub_patch_width = current_patch_bbox[2] - i * width_stride if i == num_of_width_patches - 1 else width_stride
            sub_patch_height = current_patch_bbox[3] - j * height_stride if j == num_of_height_patches - 1 else height_stride
            sub_patches.append([current_patch_bbox[0] + i * width_stride, current_patch_bbox[1] + j * height_stride,
                                sub_patch_width, sub_patch_height])
    return sub_patches, width_stride, height_stride


def iou(bbox1, bbox2):
    x1 = max(bbox1[0], bbox2[0])
    y1 = max(bbox1[1], bbox2[1])
    x2 = min(bbox1[0] + bbox1[2], bbox2[0] + bbox2[2])
    y2 = min(bbox1[1] + bbox1[3], bbox2[1] + bbox2[3])
    inter_area = max(0, x2 - x1) * max(0, y2 - y1)
    return inter_area / (bbox1[2] * bbox1[3] + bbox2[2] * bbox2[3] - inter_area)


@functools.total_ordering
class Prioritize:
    # visual_search.py:378-389: compares priority only; ties are resolved by heapq's sift order
    def __init__(self, priority, item):
        self.priority = priority
        self.item = item

    def __eq__(self, other):
        return self.priority == other.priority

    def __lt__(self, other):
        return self.priority < other.priority


# ---------------------------------------------------------------------------------------------------------------
# heat-map scoring on the GPU (visual_search.py:255-275, :420-427, :453-462)
# ---------------------------------------------------------------------------------------------------------------
class Heatmap:
    """A clamped full-resolution target-cue map resident on the GPU with its (max, min, sum)."""
    __slots__ = ("map", "stats", "h", "w", "_stats_host")

    def __init__(self, map_, stats, h, w):
        self.map, self.stats, self.h, self.w = map_, stats, h, w
        self._stats_host = None

    def host_stats(self):
        if self._stats_host is None:
            self._stats_host = self.stats.cpu().numpy()
        return self._stats_host

    def normalized(self):
        """normalize_score(...) as a [h,w,1] fp32 numpy array (what the reference stores in search_path)"""
        mx, mn = float(self.host_stats()[0]), float(self.host_stats()[1])
        m = self.map.cpu().numpy()
        m = (m - np.float32(mn)) / np.float32(mx - mn) if mx != mn else m * 0
        return m.reshape(self.h, self.w, 1)

    def __array__(self, dtype=None, copy=None):
        a = self.normalized()
        return a.astype(dtype) if dtype is not None else a


class CudaScorer:
    """Sub-patch mass of (ancestor) heat maps with the sm_100a kernels; one D2H of a few scalars per expansion."""

    def from_low_res(self, low_res, h, w):
        hm, stats = ops.heatmap(low_res.contiguous(), h, w, clamp=True, with_stats=True)
        return Heatmap(hm, stats, h, w)

    def from_full_res(self, tensor, h, w):
        """a vsm object that already returns the H x W map (generic VSM.inference API): upload if needed, then reduce"""
        t = tensor.reshape(h, w)
        if not t.is_cuda:
            t = t.cuda()
        t = t.float().contiguous()
        # identity resample (LH=h, LW=w) reuses the heat-map kernel for clamp + (max,min,sum)
        hm, stats = ops.heatmap(t, h, w, clamp=True, with_stats=True)
        return Heatmap(hm, stats, h, w)

    def rect_sums(self, jobs):
        """jobs: list of (Heatmap, rects int [n,4]); returns list of float64 arrays; a single sync."""
        outs = []
        for hmap, rects in jobs:
            r = torch.tensor(np.asarray(rects, dtype=np.int32), device=hmap.map.device)
            outs.append(ops.rect_sums(hmap.map, r, hmap.stats))
        cat = torch.cat(outs).cpu().numpy()
        res, o = [], 0
        for _, rects in jobs:
            res.append(cat[o:o + len(rects)])
            o += len(rects)
        return res


def _sub_scores_from_sums(sums, n_children):
    """get_subpatch_scores (visual_search.py:255-266) given rectangle sums of the normalised map: sums[:n] children,
    sums[n] whole patch.  Arithmetic in numpy float32 like the reference's `.sum()` results."""
    total = np.float32(sums[n_children])
    out = []
    for i in range(n_children):
        s = np.float32(sums[i])
        if total > 0:
            s = np.float32(s / total)
        else:
            s = np.float32(s * 0)
        out.append(s)
    return out


# ---------------------------------------------------------------------------------------------------------------
# one search as an explicit state machine
# ---------------------------------------------------------------------------------------------------------------
class _NodeEval:
    """Outputs of one detection-mode evaluation of a crop, reduced to what the controller consumes."""
    __slots__ = ("top_logit", "top_box", "n_logits", "boxes", "scores", "low_res", "full_map", "heat")

    def __init__(self):
        self.top_logit = None    # python float (sigmoid score)
        self.top_box = None      # torch CPU float32 [4] cxcywh in (0,1)
        self.n_logits = 0
        self.boxes = None        # full [P,4] (kept on device / lazily fetched; needed only for the root success case)
        self.scores = None
        self.low_res = None      # [4g,4g] fp32 device
        self.full_map = None     # H x W tensor from a generic vsm.inference
        self.heat = None         # Heatmap once materialised


class SearchState:
    def __init__(self, image, target_object_name, smallest_size, confidence_high=0.5, confidence_low=0.3,
                 target_cue_threshold=6.0, target_cue_threshold_decay=0.7, target_cue_threshold_minimum=3.0):
        self.image = image
        self.target = target_object_name
        self.smallest_size = smallest_size
        self.confidence_high, self.confidence_low = confidence_high, confidence_low
        self.thr, self.thr_decay, self.thr_min = target_cue_threshold, target_cue_threshold_decay, target_cue_threshold_minimum
        init_patch = dict(bbox=[0, 0, image.width, image.height], scale_level=1, score=None, parent_index=-1)
        self.search_path = [init_patch]
        self.queue = PriorityQueue()
        self.current = init_patch
        self.cache = {}               # bbox tuple -> _NodeEval
        self.done = False
        self.success = False
        self.all_valid_boxes = None
        self.n_evals = 0

    # -- frontier ------------------------------------------------------------------------------------------
    @staticmethod
    def key(patch):
        return tuple(int(v) for v in patch["bbox"])

    def wanted(self, k):
        """Nodes whose evaluation is (or is likely to be) needed next: the current node first, then the best k-1 queue
        entries (speculation; does not touch the heap order)."""
        out = []
        if not self.done and self.key(self.current) not in self.cache:
            out.append(self.current)
        if k > len(out) and not self.done:
            rest = sorted(self.queue.queue)
            for pr in rest:
                if len(out) >= k:
                    break
                if self.key(pr.item) not in self.cache and all(pr.item is not o for o in out):
                    out.append(pr.item)
        return out

    def crop(self, patch):
        b = patch["bbox"]
        return self.image.crop((int(b[0]), int(b[1]), int(b[0] + b[2]), int(b[1] + b[3])))


class SearchController:
    """Runs one or many SearchStates against a vsm object, batching node evaluations."""

    def __init__(self, vsm, scorer=None, batch_size=1, extract_noun_chunks=None):
        self.vsm = vsm
        self.scorer = scorer if scorer is not None else CudaScorer()
        self.batch_size = max(1, int(batch_size))
        self.extract_noun_chunks = extract_noun_chunks
        self.regions = hasattr(vsm, "detect_regions")
        self.batched = self.regions or hasattr(vsm, "detect_batch")

    # -- evaluation ----------------------------------------------------------------------------------------
    def _evaluate(self, requests):
        """requests: list of (state, patch).  Fills state.cache."""
        if not requests:
            return
        if self.batched:
            questions = [DETECTION_QUESTION.format(st.target) for st, p in requests]
            if self.regions:       # the crop is cut on the GPU from the resident search image
                results = self.vsm.detect_regions([(st.image, p["bbox"]) for st, p in requests], questions)
            else:
                results = self.vsm.detect_batch([st.crop(p) for st, p in requests], questions)
            for (st, p), ev in zip(requests, results):
                st.cache[st.key(p)] = ev
                st.n_evals += 1
            return
        import copy
        for st, p in requests:
            patch_img = st.crop(p)
            boxes, logits, hm = self.vsm.inference(copy.deepcopy(patch_img), DETECTION_QUESTION.format(st.target), mode="detection")
            ev = _NodeEval()
            ev.n_logits = len(logits)
            if ev.n_logits > 0:
                flat = logits.view(-1)
                ti = int(flat.argmax())
                ev.top_logit = flat.max()
                ev.top_box = boxes[ti].view(4).clone()
                ev.boxes, ev.scores = boxes, logits
            ev.full_map = hm
            st.cache[st.key(p)] = ev
            st.n_evals += 1

    # -- one node, exactly visual_search_queue's body (visual_search.py:390-473) -----------------------------
    def _process(self, st: SearchState):
        cur = st.current
        bb = cur["bbox"]
        lvl = cur["scale_level"]
        ev = st.cache[st.key(cur)]
        pw, ph = int(bb[0] + bb[2]) - int(bb[0]), int(bb[1] + bb[3]) - int(bb[1])
        if ev.n_logits > 0:
            top_logit = ev.top_logit
            final_bbox = ev.top_box.view(4) * torch.Tensor([pw, ph, pw, ph])
            final_bbox[:2] -= final_bbox[2:] / 2
            if top_logit > st.confidence_high:
                st.search_path[-1]["detection_result"] = final_bbox
                if len(st.search_path) == 1:
                    boxes, scores = ev.boxes, ev.scores
                    if boxes.is_cuda:
                        boxes, scores = boxes.cpu(), scores.cpu()
                    av = boxes[scores.view(-1) > 0.5].view(-1, 4)
                    av = av * torch.Tensor([[pw, ph, pw, ph]])
                    av[:, :2] -= av[:, 2:] / 2
                    st.all_valid_boxes = av
                return True
            st.search_path[-1]["temp_detection_result"] = (top_logit, final_bbox)
        if min(bb[2], bb[3]) <= st.smallest_size:
            return False
        h, w = bb[3], bb[2]
        if ev.heat is None:
            ev.heat = self.scorer.from_low_res(ev.low_res, h, w) if ev.low_res is not None else self.scorer.from_full_res(ev.full_map, h, w)
        subs, _, _ = get_sub_patches(bb, *split_4subpatches(bb))
        idx = len(st.search_path) - 1

        def jobs_for(heat_of_cur):
            jobs = []
            tmp = cur
            while True:
                hm_t = heat_of_cur if tmp is cur else tmp["_heat"]
                tb = tmp["bbox"]
                rects = [[s[0] - tb[0], s[1] - tb[1], s[2], s[3]] for s in subs] + [[0, 0, tb[2], tb[3]]]
                jobs.append((hm_t, rects, tmp["scale_level"]))
                if tmp["parent_index"] == -1:
                    break
                tmp = st.search_path[tmp["parent_index"]]
            return jobs

        # optimistic strong-cue path: statistics and all rectangle sums come back in ONE device->host copy
        jobs = jobs_for(ev.heat)
        sums = self.scorer.rect_sums([(j[0], j[1]) for j in jobs])
        score_max = float(ev.heat.host_stats()[0])
        threshold = max(st.thr_min, st.thr * (st.thr_decay) ** (lvl - 1))
        final_heat = ev.heat
        if not (score_max > threshold):
            # weak cue: ask the VSM where the object would be, then segment that region (visual_search.py:427-443)
            import copy
            patch_img = st.crop(cur)
            vqa_results = self.vsm.inference(copy.deepcopy(patch_img), CUE_QUESTION.format(st.target), mode="vqa")
            phrase = vqa_results.split("most likely to appear")[-1].strip()
            if phrase.endswith("."):
                phrase = phrase[:-1]
            phrase = phrase.split(st.target)[-1]
            chunks = self.extract_noun_chunks(phrase) if self.extract_noun_chunks is not None else []
            phrase = chunks[0] if len(chunks) == 1 else "region {}".format(phrase)
            cue = self.vsm.inference(copy.deepcopy(patch_img), DETECTION_QUESTION.format(phrase), mode="segmentation")
            final_heat = cue if isinstance(cue, Heatmap) else self.scorer.from_full_res(cue, h, w)
            st.search_path[idx]["context_cue"] = vqa_results + "#" + phrase
            jobs = jobs_for(final_heat)
            sums = self.scorer.rect_sums([(j[0], j[1]) for j in jobs])
        st.search_path[idx]["_heat"] = final_heat
        st.search_path[idx]["final_heatmap"] = final_heat      # lazily converts to the reference's [h,w,1] numpy array
        basic = [0] * len(subs)
        for (hm_t, rects, level), sm in zip(jobs, sums):
            tmp_scores = _sub_scores_from_sums(sm, len(subs))
            basic = [basic[i] + tmp_scores[i] / (4 ** level) for i in range(len(subs))]
        for sp, sc in zip(subs, basic):
            info = dict(bbox=sp, scale_level=lvl + 1, score=sc, parent_index=idx)
            st.queue.put(Prioritize(-info["score"], info))
        return False

    def _advance(self, st: SearchState):
        """Commit as many nodes as the cache allows, in the reference's pop order."""
        while not st.done:
            if st.key(st.current) not in st.cache:
                return
            ok = self._process(st)
            if ok:
                st.done, st.success = True, True
                return
            if st.queue.empty():
                st.done = True
                return
            st.current = st.queue.get().item
            st.search_path.append(st.current)

    def run(self, states):
        while True:
            for st in states:
                self._advance(st)
            active = [st for st in states if not st.done]
            if not active:
                break
            # mandatory nodes first, then fill the batch with speculation, round-robin over the active searches
            reqs = [(st, st.current) for st in active]
            room = self.batch_size - len(reqs)
            if room > 0:
                per = max(1, room // len(active)) + 1
                extra = []
                for st in active:
                    for p in st.wanted(per)[1:]:
                        extra.append((st, p))
                reqs += extra[:room]
            self._evaluate(reqs)
        return [self._finish(st) for st in states]

    @staticmethod
    def _finish(st: SearchState):
        # visual_search.py:496-516
        path_length = len(st.search_path)
        final_step = st.search_path[-1]
        ok = st.success
        if not ok:
            max_logit, final_step, path_length = 0, None, 0
            for i, step in enumerate(st.search_path):
                if "temp_detection_result" in step and step["temp_detection_result"][0] > max_logit:
                    max_logit, final_step, path_length = step["temp_detection_result"][0], step, i + 1
            if final_step is None:
                raise RuntimeError("no detection result on any search step (the reference dereferences None here: "
                                   "visual_search.py:509)")
            final_step["detection_result"] = final_step["temp_detection_result"][1]
            if max_logit >= st.confidence_low:
                ok = True
        return final_step, path_length, ok, st.all_valid_boxes


def visual_search(vsm, image, target_object_name, target_bbox, smallest_size, confidence_high=0.5, confidence_low=0.3,
                  target_cue_threshold=6.0, target_cue_threshold_decay=0.7, target_cue_threshold_minimum=3.0,
                  visualize=False, save_path=None, scorer=None, batch_size=None, extract_noun_chunks=None,
                  return_state=False):
    """Same contract as the reference's visual_search() (visual_search.py:484-516)."""
    if visualize:
        raise NotImplementedError("visualisation (cv2 overlays, visual_search.py:339-376) is outside the hot path")
    st = SearchState(image, target_object_name, smallest_size, confidence_high, confidence_low, target_cue_threshold,
                     target_cue_threshold_decay, target_cue_threshold_minimum)
    if batch_size is None:
        batch_size = getattr(vsm, "frontier_batch", 1)
    ctl = SearchController(vsm, scorer, batch_size, extract_noun_chunks)
    res = ctl.run([st])[0]
    return res + (st,) if return_state else res


def visual_search_many(vsm, jobs, batch_size=8, scorer=None, extract_noun_chunks=None, **kw):
    """Run several independent searches in lock-step so their frontiers share GPU batches.
    jobs: list of (image, target_object_name, smallest_size).  Returns (results, states)."""
    states = [SearchState(img, name, ss, **kw) for img, name, ss in jobs]
    ctl = SearchController(vsm, scorer, batch_size, extract_noun_chunks)
    return ctl.run(states), states
