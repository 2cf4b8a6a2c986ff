"""torchrun --nproc-per-node N tools/multi_gpu_check.py : sharded frontier with the real CUDA engine (tiny config, NCCL)
must reproduce the single-GPU search on every rank."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from PIL import Image


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vstar_b200 import synth
    from vstar_b200.config import VSMConfig
    from vstar_b200.engine import VSMEngine, VSMWeights
    from vstar_b200.sharded import ShardedVSM
    from vstar_b200.visual_search import visual_search
    from vstar_b200.vsm import VSM
    j = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "tiny_config.json")))
    cfg = VSMConfig(**j["cfg"])
    sd = synth.synthetic_state_dict(cfg, seed=j["weight_seed"])
    eng = VSMEngine(VSMWeights.from_state_dict(cfg, sd, device=f"cuda:{local}"))
    prompt, ans = synth.synthetic_prompt(cfg, n_text=24, seed=5)

    class V(VSM):
        def _ids(self, q):
            return prompt[0].tolist()

    vsm = V(engine=eng, forced_answer_ids=ans.tolist(), frontier_batch=4)
    img = Image.fromarray(np.random.default_rng(31).integers(0, 256, (512, 640, 3), dtype=np.uint8), "RGB")
    kw = dict(confidence_high=2.0, target_cue_threshold=-1e9, target_cue_threshold_minimum=-1e9)
    fs, pl, ok, av, st = visual_search(vsm, img, "mug", None, 130, return_state=True, **kw)
    sh = ShardedVSM(vsm, device=f"cuda:{local}")
    fs2, pl2, ok2, av2, st2 = visual_search(sh, img, "mug", None, 130, return_state=True, batch_size=4 * world, **kw)
    t1 = [tuple(s["bbox"]) for s in st.search_path]
    t2 = [tuple(s["bbox"]) for s in st2.search_path]
    same = (t1 == t2) and pl == pl2 and bool(torch.equal(fs["detection_result"], fs2["detection_result"]))
    # context-cue branch on every expandable node: cue answers travel as strings (all_gather_object), cue maps as crop records
    from vstar_b200 import noun_chunks
    noun_chunks.set_nlp(lambda text: [])
    kw2 = dict(confidence_high=2.0, target_cue_threshold=1e9, target_cue_threshold_minimum=1e9)
    a = visual_search(vsm, img, "mug", None, 200, return_state=True, **kw2)[4]
    b = visual_search(sh, img, "mug", None, 200, return_state=True, batch_size=4 * world, **kw2)[4]
    same_cue = [tuple(s["bbox"]) for s in a.search_path] == [tuple(s["bbox"]) for s in b.search_path] and \
        [s.get("context_cue") for s in a.search_path] == [s.get("context_cue") for s in b.search_path] and \
        [s["score"] for s in a.search_path[1:]] == [s["score"] for s in b.search_path[1:]] and any("context_cue" in s for s in b.search_path)
    same = same and same_cue
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(world=world, nodes=len(t1), sharded_equals_single=bool(flag.item()), gathered_bytes=sh.gathered_bytes)))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
