"""Frontier sharding over the GPUs of one box: SPMD control, data-parallel crop evaluation.

Every rank runs the SAME search controller on the SAME search images (deterministic host code), so no decision ever
has to be broadcast.  Only `detect_regions` is distributed: the frontier batch is dealt round-robin over the ranks, each
rank evaluates its share with its own replica of the weights, and one fixed-size record per crop — top score, top box,
all 2304 (score, box) rows and the 192x192 low-res mask (194 KB) — is all-gathered (NCCL over NVLink / NVSwitch; gloo in
the CPU tests) so that every rank holds every result and commits nodes in the reference's pop order.  This is the only
exchange step on the path; there is no reduction and no tensor parallelism.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .visual_search import _NodeEval


class ShardedVSM:
    def __init__(self, vsm, group=None, device=None):
        self.vsm = vsm
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.frontier_batch = getattr(vsm, "frontier_batch", 8) * self.world
        self.gathered_bytes = 0

    # everything that is not the batched detection goes to the local replica (vqa / segmentation of the weak-cue branch
    # are pure functions of the crop, so every rank computes the same value)
    def inference(self, image, question, mode="segmentation"):
        return self.vsm.inference(image, question, mode)

    @staticmethod
    def _pack(ev, P, L):
        rec = torch.zeros(6 + 5 * P + L, dtype=torch.float32, device=ev.low_res.device)
        rec[0] = float(ev.top_logit)
        rec[1:5] = ev.top_box.to(rec.device)
        rec[5] = float(ev.n_logits)
        rec[6:6 + P] = ev.scores.reshape(-1).to(rec.device)
        rec[6 + P:6 + 5 * P] = ev.boxes.reshape(-1).to(rec.device)
        rec[6 + 5 * P:] = ev.low_res.reshape(-1)
        return rec

    @staticmethod
    def _unpack(rec, P, side):
        ev = _NodeEval()
        host = rec[:6].cpu()
        ev.top_logit = float(host[0])
        ev.top_box = host[1:5].clone()
        ev.n_logits = int(host[5])
        ev.scores = rec[6:6 + P].view(P, 1)
        ev.boxes = rec[6 + P:6 + 5 * P].view(P, 4)
        ev.low_res = rec[6 + 5 * P:].view(side, side)
        return ev

    def detect_regions(self, regions, questions):
        n = len(regions)
        w, r = self.world, self.rank
        mine = list(range(r, n, w))
        evs = self.vsm.detect_regions([regions[i] for i in mine], [questions[i] for i in mine]) if mine else []
        n_max = (n + w - 1) // w
        # record geometry from the first local result; ranks without work learn it from the gathered header
        if evs:
            P = evs[0].scores.numel()
            side = evs[0].low_res.shape[-1]
            dev = evs[0].low_res.device
        else:
            P, side, dev = 0, 0, self.device
        geo = torch.tensor([P, side], dtype=torch.int64, device=dev)
        dist.all_reduce(geo, op=dist.ReduceOp.MAX, group=self.group)
        P, side = int(geo[0]), int(geo[1])
        L = side * side
        mine_rec = torch.zeros((n_max, 6 + 5 * P + L), dtype=torch.float32, device=dev)
        for k, ev in enumerate(evs):
            mine_rec[k] = self._pack(ev, P, L)
        parts = [torch.empty_like(mine_rec) for _ in range(w)]
        dist.all_gather(parts, mine_rec, group=self.group)
        self.gathered_bytes += mine_rec.numel() * 4 * w
        out = []
        for i in range(n):
            out.append(self._unpack(parts[i % w][i // w], P, side))
        return out


def broadcast_loader(get, src=0, group=None, device="cuda"):
    """Weight broadcast at start-up: only rank `src` reads the checkpoint; every rank builds its replica
    through the returned `name -> tensor` callable, which on `src` loads the tensor and broadcasts it (NCCL over NVLink on
    the GPUs, gloo on CPU) and elsewhere receives it.  All ranks must request the same names in the same order - which they
    do, because `CoreWeights` / `VSMWeights` / `VQAWeights` walk the reference's key layout deterministically.
    `get` may be None on the other ranks."""
    rank = dist.get_rank(group)

    def load(name):
        if rank == src:
            t = get(name).to(device)
            meta = [(tuple(t.shape), t.dtype)]
        else:
            t, meta = None, [None]
        dist.broadcast_object_list(meta, src=src, group=group)
        shape, dtype = meta[0]
        if rank != src:
            t = torch.empty(shape, dtype=dtype, device=device)
        t = t.contiguous()
        dist.broadcast(t, src=src, group=group)
        return t

    return load
