"""N > 1 path on CPU: world_size-2 gloo.  The sharded frontier (vstar_b200/sharded.py) must give every rank exactly the
results of the un-sharded evaluation, and the SPMD controller must then walk the reference's trajectory on both ranks."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import NumpyScorer, StubVSM, synth_image

G = os.path.join(os.path.dirname(__file__), "golden")


class RegionStub(StubVSM):
    """pure function of the crop, through the detect_regions interface the CUDA VSM exposes (CPU tensors here)"""

    def __init__(self):
        super().__init__()
        self.n_local = 0

    def detect_regions(self, regions, questions):
        from vstar_b200.visual_search import _NodeEval
        out = []
        for src, b in regions:
            self.n_local += 1
            im = src.crop((int(b[0]), int(b[1]), int(b[0] + b[2]), int(b[1] + b[3])))
            boxes, logits, hm = StubVSM.inference(self, im, "", "detection")
            ev = _NodeEval()
            ev.n_logits = len(logits)
            ev.top_logit = float(logits.view(-1).max())
            ev.top_box = boxes[int(logits.view(-1).argmax())].clone()
            ev.boxes, ev.scores = boxes, logits
            # a 192x192 "low-res" map; the controller up-samples it like the real one
            arr = np.asarray(im, dtype=np.uint8)
            s = int(arr[::max(1, arr.shape[0] // 16), ::max(1, arr.shape[1] // 16)].astype(np.int64).sum()) % (2 ** 31)
            ev.low_res = torch.from_numpy(np.random.default_rng(s).standard_normal((192, 192)).astype(np.float32) * 3)
            out.append(ev)
        return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vstar_b200 import visual_search as VS
    from vstar_b200.sharded import ShardedVSM
    img = synth_image(21, 512, 512)
    kw = dict(confidence_high=2.0, target_cue_threshold=-1e9, target_cue_threshold_minimum=-1e9)
    inner = RegionStub()
    sh = ShardedVSM(inner, device="cpu")
    fs, pl, ok, av, st = VS.visual_search(sh, img, "mug", None, 100, scorer=NumpyScorer(), batch_size=8, return_state=True, **kw)
    traj = [tuple(s["bbox"]) for s in st.search_path]
    # un-sharded run in the same process
    ref = RegionStub()
    fs2, pl2, ok2, av2, st2 = VS.visual_search(ref, img, "mug", None, 100, scorer=NumpyScorer(), batch_size=8, return_state=True, **kw)
    q.put((rank, traj == [tuple(s["bbox"]) for s in st2.search_path], pl == pl2, inner.n_local, ref.n_local,
           torch.equal(fs["detection_result"], fs2["detection_result"])))
    dist.destroy_process_group()


def test_sharded_frontier_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    locals_ = []
    for rank, same_traj, same_pl, n_local, n_ref, same_det in res:
        assert same_traj and same_pl and same_det, (rank, same_traj, same_pl, same_det)
        locals_.append(n_local)
        total = n_ref
    # the two ranks split the evaluations between them (speculative batches make the split uneven but complete)
    assert sum(locals_) >= total and max(locals_) < total


def _bcast_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vstar_b200 import synth
    from vstar_b200.config import tiny_config
    from vstar_b200.sharded import broadcast_loader
    cfg = tiny_config()
    shapes = synth.state_dict_shapes(cfg)
    names = sorted(shapes)[:40]
    # only rank 0 "has the checkpoint"; rank 1 passes no loader at all
    get = (lambda n: synth.synthetic_tensor(n, shapes[n], seed=77)) if rank == 0 else None
    load = broadcast_loader(get, src=0, device="cpu")
    got = {n: load(n) for n in names}
    ok = all(torch.equal(got[n], synth.synthetic_tensor(n, shapes[n], seed=77)) for n in names)
    q.put((rank, ok, len(got)))
    dist.destroy_process_group()


def test_weight_broadcast_world2_gloo():
    """start-up weight broadcast: rank 0 reads, every rank ends up with identical replicas"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] == 40 for r in res)
