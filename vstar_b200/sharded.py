"""Frontier sharding over the GPUs of one box: SPMD control, data-parallel crop evaluation.

Every rank runs the SAME search controller on the SAME search images (deterministic host code), so no decision ever has to
be broadcast.  Only the batched crop evaluation is distributed: a frontier batch is dealt round-robin over the ranks, each
rank evaluates its share with its own replica of the weights and produces one fixed-size CROP RECORD per crop on the device
(records.py: best score/box, boxes above 0.5, heat-map statistics, rectangle sums of the normalised map over the crop's
quad-tree descendants; 76 + #rectangles floats ~ 0.3-1.7 KB, 5.5 KB for the root of an 8192^2 / 256 search).  ONE
`all_gather_into_tensor` of those records per batch (NCCL over NVLink / NVSwitch; gloo in the CPU tests) gives every rank
every result; nodes are then committed in the reference's pop order on all ranks alike.  The 192 x 192 masks stay on the rank
that computed them (the controller never needs them: the rectangle sums are in the record); `search_path[i]['final_heatmap']`
of a node owned by another rank is fetched by a broadcast only if somebody materialises it (visualisation).
There is no reduction, no tensor parallelism and no per-record host synchronisation on this path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .records import pyramid_rects, record_floats
from .visual_search import _NodeEval


class ShardedVSM:
    """Wraps a local VSM replica (anything with detect_regions_launch / detect_regions_finish producing record tensors)."""

    def __init__(self, vsm, group=None, device=None):
        self.vsm = vsm
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device if device is not None else getattr(getattr(vsm, "engine", None), "dev", "cpu")
        self.frontier_batch = getattr(vsm, "frontier_batch", 8) * self.world
        self.gathered_bytes = 0          # bytes this rank received through the record all-gathers
        self.gathers = 0
        self._bufs = {}
        self._d2h_stream = None

    # everything that is not the batched detection goes to the local replica (vqa / segmentation of the weak-cue branch
    # are pure functions of the crop, so every rank computes the same value)
    def inference(self, image, question, mode="segmentation"):
        return self.vsm.inference(image, question, mode)

    cue_records = True

    def inference_many(self, regions, questions, mode, smallest_sizes=None):
        """weak-cue calls of all parked searches, dealt over the ranks like a frontier batch: cue answers ('vqa') come back as
        strings through one all_gather_object, cue segmentations as crop records through the same all-gather as detections"""
        n, w, r = len(regions), self.world, self.rank
        if mode == "segmentation" and smallest_sizes is not None and getattr(self.vsm, "cue_records", False):
            return self.detect_regions_finish(self.detect_regions_launch(regions, questions, smallest_sizes, mode="segmentation"))
        mine = list(range(r, n, w))
        if hasattr(self.vsm, "inference_many"):
            local = self.vsm.inference_many([regions[i] for i in mine], [questions[i] for i in mine], mode) if mine else []
        else:
            local = [self.vsm.inference(regions[i][0].crop((int(regions[i][1][0]), int(regions[i][1][1]), int(regions[i][1][0] + regions[i][1][2]),
                                                            int(regions[i][1][1] + regions[i][1][3]))), questions[i], mode) for i in mine]
        if mode != "vqa":
            # maps cannot travel as objects: without the record interface every rank evaluates every cue segmentation itself
            if hasattr(self.vsm, "inference_many"):
                return self.vsm.inference_many(regions, questions, mode)
            return [self.vsm.inference(src.crop((int(b[0]), int(b[1]), int(b[0] + b[2]), int(b[1] + b[3]))), q, mode) for (src, b), q in zip(regions, questions)]
        parts = [None] * w
        dist.all_gather_object(parts, local, group=self.group)
        return [parts[i % w][i // w] for i in range(n)]

    def _buffer(self, name, shape, dtype, pinned=False):
        key = (name, tuple(shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            if pinned:
                b = torch.empty(shape, dtype=dtype).pin_memory()
            else:
                b = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = b
        return b

    def detect_regions_launch(self, regions, questions, smallest_sizes, mode="detection"):
        n, w, r = len(regions), self.world, self.rank
        mine = list(range(r, n, w))
        n_max = (n + w - 1) // w
        # the record length is a function of the batch's geometry only, so every rank computes the same value
        R = record_floats(max(len(pyramid_rects(b, ss)) for (_, b), ss in zip(regions, smallest_sizes)))
        kw = {} if mode == "detection" else {"mode": mode}
        local = self.vsm.detect_regions_launch([regions[i] for i in mine], [questions[i] for i in mine],
                                               [smallest_sizes[i] for i in mine], rec_len=R, **kw) if mine else None
        cuda = torch.device(self.device).type == "cuda"
        # double-buffered by launch parity: with two batches in flight the previous gather's buffers are still being read
        slot = self.gathers % 4
        send = self._buffer(("send", slot), (n_max, R), torch.float32)
        recv = self._buffer(("recv", slot), (w * n_max, R), torch.float32)
        if local is not None:
            rec = local["rec"]
            send[:rec.shape[0]].copy_(rec)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        self.gathers += 1
        self.gathered_bytes += recv.numel() * 4
        pend = dict(n=n, mine=mine, local=local, regions=regions, smallest=list(smallest_sizes), recv=recv, n_max=n_max, R=R, done=None)
        if cuda:
            main = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)
            if self._d2h_stream is None:
                self._d2h_stream = torch.cuda.Stream(device=self.device)
            host = self._buffer(("host", slot), (w * n_max, R), torch.float32, pinned=True)
            with torch.cuda.stream(self._d2h_stream):
                self._d2h_stream.wait_event(ready)
                host.copy_(recv, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._d2h_stream)
            pend["host"], pend["done"] = host, done
        else:
            pend["host"] = recv
        return pend

    def detect_regions_finish(self, pend):
        if pend["done"] is not None:
            pend["done"].synchronize()
        rows = pend["host"].numpy()
        w, n_max = self.world, pend["n_max"]
        own = {}
        if pend["local"] is not None:
            # local finish: resolves this rank's draft verification and hands back the device-side tensors of its own crops
            for i, ev in zip(pend["mine"], self.vsm.detect_regions_finish(pend["local"])):
                own[i] = ev
        out = []
        for i in range(pend["n"]):
            ev = _NodeEval.from_record(rows[(i % w) * n_max + i // w], pend["regions"][i][1], pend["smallest"][i])
            src = i % w
            mine = own.get(i)
            # masks / detections stay with their owner; the two fetchers are COLLECTIVE (every rank calls them at the same
            # point of the same controller), so ev.low_res is left unset even on the owning rank
            ev.fetch_valid = self._valid_fetcher(mine, src)
            ev.fetch_low_res = self._low_res_fetcher(mine, src)
            out.append(ev)
        return out

    def detect_regions(self, regions, questions, smallest_sizes=None):
        if smallest_sizes is None:
            smallest_sizes = [max(1, min(int(b[2]), int(b[3])) // 2) for _, b in regions]
        return self.detect_regions_finish(self.detect_regions_launch(regions, questions, smallest_sizes))

    # -- rare follow-ups (SPMD: every rank reaches them at the same point of the same controller) ---------------------
    def _valid_fetcher(self, mine, src):
        """more than 16 boxes above 0.5 at a successful ROOT (visual_search.py:406-410): the owner broadcasts the full list"""

        def fetch():
            n = torch.zeros(1, dtype=torch.int64, device=self.device)
            if mine is not None:
                boxes = mine.fetch_valid().to(self.device).float().contiguous()
                n[0] = boxes.shape[0]
            dist.broadcast(n, src=src, group=self.group)
            if mine is None:
                boxes = torch.empty((int(n[0]), 4), dtype=torch.float32, device=self.device)
            dist.broadcast(boxes, src=src, group=self.group)
            return boxes.cpu()

        return fetch

    def _low_res_fetcher(self, mine, src):
        def fetch():
            shape = torch.zeros(2, dtype=torch.int64, device=self.device)
            if mine is not None:
                low = mine.low_res.to(self.device).float().contiguous()
                shape[0], shape[1] = low.shape[-2], low.shape[-1]
            dist.broadcast(shape, src=src, group=self.group)
            if mine is None:
                low = torch.empty((int(shape[0]), int(shape[1])), dtype=torch.float32, device=self.device)
            dist.broadcast(low, src=src, group=self.group)
            return low

        return fetch


def broadcast_loader(get, src=0, group=None, device="cuda", index=None):
    """Weight broadcast at start-up: only rank `src` reads the checkpoint; every rank builds its replica
    through the returned `name -> tensor` callable, which on `src` loads the tensor and broadcasts it (NCCL over NVLink on
    the GPUs, gloo on CPU) and elsewhere receives it.  All ranks must request the same names in the same order - which they
    do, because `CoreWeights` / `VSMWeights` / `VQAWeights` walk the reference's key layout deterministically.
    `get` may be None on the other ranks.

    index: {name: (shape, dtype)} known on `src` (a `SafetensorsStream` carries one in `.index`; bench.py passes its synthetic
    shape table).  It is broadcast ONCE here, so the per-tensor path is a single `dist.broadcast` with no object exchange;
    without an index every tensor's shape/dtype is sent first (one small object broadcast per tensor).
    `load.stats` counts tensors / bytes."""
    rank = dist.get_rank(group)
    stats = dict(tensors=0, bytes=0, seconds=0.0)
    if rank == src and index is None and hasattr(get, "index"):
        index = {k: (tuple(v[2]), v[1]) for k, v in get.index.items()}          # SafetensorsStream: (file, dtype, shape, begin, end)
    box = [index if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    index = box[0]

    def load(name):
        if rank == src:
            t = get(name).to(device)
        if index is not None and name in index:
            shape, dtype = index[name]
            if rank == src:
                assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, (name, t.shape, shape, t.dtype, dtype)
        else:
            meta = [(tuple(t.shape), t.dtype)] if rank == src else [None]
            dist.broadcast_object_list(meta, src=src, group=group)
            shape, dtype = meta[0]
        if rank != src:
            t = torch.empty(shape, dtype=dtype, device=device)
        t = t.contiguous()
        dist.broadcast(t, src=src, group=group)
        stats["tensors"] += 1
        stats["bytes"] += t.numel() * t.element_size()
        return t

    load.stats = stats
    return load
