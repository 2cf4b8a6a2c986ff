"""The product's iterative / batched search controller must expand nodes in exactly the reference's order.
Pinned against the trajectories produced by the REAL reference visual_search() (tests/golden/search_*.npz).
The heat-map sums are injected (tests/helpers.NumpyScorer) so this ordering logic is testable without a GPU;
the CUDA scorer itself is covered by the gpu tests."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import FakeNLP, NumpyScorer, RecordStub, StubVSM, synth_image
from vstar_b200 import noun_chunks
from vstar_b200 import visual_search as VS

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def fake_nlp():
    """the goldens were generated with the reference's spaCy pipeline replaced by tests.helpers.FakeNLP (oracle/make_golden.py)"""
    noun_chunks.set_nlp(FakeNLP())
    yield
    noun_chunks.set_nlp(None)


EDGE_TAGS = ["edge_root_hit", "edge_deep_hit", "edge_tiny", "edge_tiny_unsure", "edge_wide", "edge_odd"]


def golden_stub(g):
    hot = str(g["hot"]) if "hot" in g.files else "None"
    return StubVSM(None if hot == "None" else hot)


ALL_TAGS = ["stub_3lvl", "stub_default", "stub_weakcue", "stub_mixcue"] + EDGE_TAGS


def check_against_golden(g, stub, fs, pl, ok, av, st):
    if "has_all_valid" in g.files:
        assert (av is not None) == bool(int(g["has_all_valid"]))
        if av is not None:
            assert np.array_equal(av.numpy(), g["all_valid_boxes"])
    assert np.array_equal(np.array([s["bbox"] for s in st.search_path]), g["trajectory"])
    assert pl == int(g["path_length"]) and int(ok) == int(g["success"])
    assert list(fs["bbox"]) == list(g["final_bbox"])
    assert np.allclose(fs["detection_result"].numpy(), g["detection_result"], rtol=0, atol=0)
    # the context-cue strings pin the noun-chunk logic (visual_search.py:430-442): answer + "#" + phrase per weak-cue node
    assert [s.get("context_cue", "") for s in st.search_path] == json.loads(str(g["context_cues"]))


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_trajectory_matches_reference(tag):
    """trajectories, VSM call sequences and return values of the REAL reference visual_search(); the edge_* cases cover its
    termination / selection branches: confident hit at the root with several valid boxes (visual_search.py:404-410), hit
    below the root (:411), root already the smallest unit (:416-417), nothing above confidence_low (:497-512), 4x1 splits
    and sizes the 2x2 grid does not divide (:234-253)"""
    g = np.load(os.path.join(G, f"search_{tag}.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    kw = json.loads(str(g["kw"]))
    stub = golden_stub(g)
    fs, pl, ok, av, st = VS.visual_search(stub, img, "mug", None, int(g["smallest"]), scorer=NumpyScorer(), return_state=True, **kw)
    check_against_golden(g, stub, fs, pl, ok, av, st)
    assert np.array_equal(np.array(stub.calls), g["calls"])


@pytest.mark.parametrize("batch", [1, 4, 32])
@pytest.mark.parametrize("tag", ALL_TAGS)
def test_record_path_pipelined_matches_reference(tag, batch):
    """the product path: crop RECORDS (rectangle-sum pyramids instead of heat maps) + pipelined launch/finish controller with
    speculative batches must still walk the reference's trajectory and return its results"""
    g = np.load(os.path.join(G, f"search_{tag}.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    kw = json.loads(str(g["kw"]))
    hot = str(g["hot"]) if "hot" in g.files else "None"
    stub = RecordStub(None if hot == "None" else hot)
    fs, pl, ok, av, st = VS.visual_search(stub, img, "mug", None, int(g["smallest"]), scorer=NumpyScorer(), batch_size=batch,
                                          return_state=True, **kw)
    check_against_golden(g, stub, fs, pl, ok, av, st)
    assert max(stub.batches) <= batch
    # the lazily materialised final_heatmap of an expanded node is the reference's [h,w,1] fp32 normalised map
    node = st.search_path[0]
    if "final_heatmap" in node and tag == "stub_3lvl":
        a = np.asarray(node["final_heatmap"])
        assert a.shape == (int(g["h"]), int(g["w"]), 1) and float(a.max()) == 1.0 and float(a.min()) == 0.0


def test_weak_cue_calls_are_batched_across_lockstep_searches():
    """several searches whose nodes take the context-cue branch at the same time: their 'vqa' and 'segmentation' calls go out
    as batches (vsm.inference_many), and every search still walks the trajectory / cue strings of the reference"""
    g = np.load(os.path.join(G, "search_stub_mixcue.npz"))
    g2 = np.load(os.path.join(G, "search_stub_weakcue.npz"))
    kw = json.loads(str(g["kw"]))
    assert kw["confidence_high"] == json.loads(str(g2["kw"]))["confidence_high"]
    jobs = [(synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"])), "mug", int(g["smallest"])) for _ in range(3)]
    stub = RecordStub()
    res, states = VS.visual_search_many(stub, jobs, batch_size=16, scorer=NumpyScorer(), **kw)
    for st in states:
        assert np.array_equal(np.array([s["bbox"] for s in st.search_path]), g["trajectory"])
        assert [s.get("context_cue", "") for s in st.search_path] == json.loads(str(g["context_cues"]))
    assert max(n for _, n in stub.cue_batches) == 3 and {k for k, _ in stub.cue_batches} == {"vqa", "segmentation"}


def test_children_of_nodes_under_evaluation_fill_spare_batch_capacity():
    """a single search with a wide batch: the root goes out together with its four children (their geometry is known before the
    root commits), so the 1 + 4 + 16 + 64 tree needs 3 GPU rounds instead of 4 - and walks the same trajectory"""
    g = np.load(os.path.join(G, "search_stub_3lvl.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    kw = json.loads(str(g["kw"]))
    runs = {}
    for spec in (True, False):
        stub = RecordStub()
        st = VS.SearchState(img, "mug", int(g["smallest"]), **kw)
        ctl = VS.SearchController(stub, NumpyScorer(), 128, depth=1, speculate_children=spec)
        ctl.run([st])
        assert np.array_equal(np.array([s["bbox"] for s in st.search_path]), g["trajectory"])
        runs[spec] = list(stub.batches)
    assert runs[True][0] == 5 and runs[False][0] == 1 and len(runs[True]) < len(runs[False])


def test_more_than_16_valid_boxes_at_the_root():
    """all_valid_boxes when the record's 16 slots overflow: fetched from the owner, equal to the map-based path"""
    img = synth_image(24, 1280, 960)
    a = VS.visual_search(StubVSM("many"), img, "mug", None, 224, scorer=NumpyScorer())
    b = VS.visual_search(RecordStub("many"), img, "mug", None, 224, scorer=NumpyScorer(), batch_size=4)
    assert a[3] is not None and a[3].shape == (30, 4) and torch.equal(a[3], b[3])
    assert a[1] == b[1] and a[2] == b[2] and torch.equal(a[0]["detection_result"], b[0]["detection_result"])


def test_missing_spacy_raises_instead_of_degrading():
    """the weak-cue branch must not silently fall back to 'region {phrase}' when spaCy is absent"""
    noun_chunks.set_nlp(None)
    try:
        import spacy  # noqa: F401
        pytest.skip("spaCy is installed here")
    except ImportError:
        pass
    g = np.load(os.path.join(G, "search_stub_weakcue.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    with pytest.raises(noun_chunks.NounChunkerUnavailable):
        VS.visual_search(StubVSM(), img, "mug", None, int(g["smallest"]), scorer=NumpyScorer(), **json.loads(str(g["kw"])))


def test_pyramid_rects_geometry():
    """record pyramid = the node, then the children of every expandable descendant, in BFS order; leaves tile the node"""
    from vstar_b200 import records as RC
    for bbox, ss in (([0, 0, 1024, 1024], 224), ([10, 20, 1001, 777], 251), ([0, 0, 2000, 420], 224), ([5, 5, 300, 1900], 300),
                     ([0, 0, 200, 150], 224)):
        rects = RC.pyramid_rects(bbox, ss)
        if not RC.expandable(bbox, ss):
            assert rects == []
            continue
        assert rects[0] == tuple(bbox)
        want, level = [tuple(bbox)], [bbox]
        while level:
            nxt = []
            for b in level:
                subs, _, _ = VS.get_sub_patches(b, *VS.split_4subpatches(b))
                assert sum(s[2] * s[3] for s in subs) == b[2] * b[3]          # children tile the parent exactly
                want += [tuple(s) for s in subs]
                nxt += [s for s in subs if min(s[2], s[3]) > ss]
            level = nxt
        assert rects == want and len(set(rects)) == len(rects)
        assert RC.record_floats(len(rects)) % 4 == 0 and RC.record_floats(len(rects)) >= 76 + len(rects)


class BatchStub(StubVSM):
    """same pure function, but through the batched entry point the CUDA VSM exposes"""

    def __init__(self):
        super().__init__()
        self.batches = []

    def detect_batch(self, images, questions):
        self.batches.append(len(images))
        out = []
        for im, q in zip(images, questions):
            boxes, logits, hm = StubVSM.inference(self, im, q, "detection")
            ev = VS._NodeEval()
            ev.n_logits = len(logits)
            ev.top_logit = logits.view(-1).max()
            ev.top_box = boxes[int(logits.view(-1).argmax())].clone()
            ev.boxes, ev.scores, ev.full_map = boxes, logits, hm
            out.append(ev)
        return out


@pytest.mark.parametrize("batch", [1, 4, 16])
def test_speculative_batching_keeps_order(batch):
    g = np.load(os.path.join(G, "search_stub_3lvl.npz"))
    img = synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"]))
    kw = json.loads(str(g["kw"]))
    stub = BatchStub()
    fs, pl, ok, av, st = VS.visual_search(stub, img, "mug", None, int(g["smallest"]), scorer=NumpyScorer(), batch_size=batch,
                                          return_state=True, **kw)
    assert np.array_equal(np.array([s["bbox"] for s in st.search_path]), g["trajectory"])
    assert pl == int(g["path_length"]) and list(fs["bbox"]) == list(g["final_bbox"])
    assert max(stub.batches) <= batch
    if batch > 1:
        assert max(stub.batches) > 1          # speculation actually batches


def test_many_searches_lockstep():
    jobs, want = [], []
    for tag in ["stub_3lvl", "stub_default"]:
        g = np.load(os.path.join(G, f"search_{tag}.npz"))
        kw = json.loads(str(g["kw"]))
        if tag == "stub_default":
            continue
        for k in range(3):
            jobs.append((synth_image(int(g["img_seed"]), int(g["w"]), int(g["h"])), "mug", int(g["smallest"])))
            want.append(g["trajectory"])
    stub = BatchStub()
    res, states = VS.visual_search_many(stub, jobs, batch_size=8, scorer=NumpyScorer(), confidence_high=2.0)
    for st, w in zip(states, want):
        assert np.array_equal(np.array([s["bbox"] for s in st.search_path]), w)


def test_deep_search_does_not_recurse():
    """the reference hits Python's recursion limit near 990 expanded nodes"""
    img = synth_image(5, 512, 512)
    stub = BatchStub()
    fs, pl, ok, av, st = VS.visual_search(stub, img, "mug", None, 10, scorer=NumpyScorer(), batch_size=64, confidence_high=2.0,
                                          return_state=True)
    assert len(st.search_path) == 1 + 4 + 16 + 64 + 256 + 1024 + 4096      # 512 -> 8 px leaves, > 990 expanded nodes


def test_prompt_and_tokenizer_helpers():
    from vstar_b200.config import tiny_config
    from vstar_b200.vsm import SyntheticTokenizer, build_prompt, tokenizer_image_token
    cfg = tiny_config()
    tok = SyntheticTokenizer(cfg)
    p = build_prompt("Please locate the mug in this image.")
    assert p.startswith("A chat between") and p.endswith("ASSISTANT:") and "<im_start><image><im_end>\n" in p
    ids = tokenizer_image_token(p, tok)
    assert ids[0] == 1 and ids.count(-200) == 1
    k = ids.index(-200)
    assert ids[k - 1] == cfg.vocab - 2 and ids[k + 1] == cfg.vocab - 1
    assert tok("Sure, [LOC].", add_special_tokens=False).input_ids.count(cfg.loc_token_idx) == 1
