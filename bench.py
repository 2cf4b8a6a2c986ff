#!/usr/bin/env python
"""bench.py — crops/s of the V* guided visual-search hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...          # the reference's CPU path (oracle port) on the host cores

A "step" = one pass of the hot path over one batch of synthetic input: `--searches` (32) independent guided visual
searches (BASELINE.json configs[1]: 1024x1024 synthetic images, smallest_size 512 => root + 4 crops each, depth 2,
bf16), run in lock-step so their crop frontiers share GPU batches; every crop evaluation = CLIP ViT-L/14 -> projector
-> Vicuna-7B-shaped prefill (draft-verified answer) -> OWL-ViT-B/16 -> SAM prompt/mask decoder -> OWL heads -> heat-map
statistics, with random-init weights of the reference architecture.

  value : crops/s with the search images already resident in HBM (uint8) when the timed region starts
  e2e   : crops/s through the public API (`visual_search_many(VSM, PIL images)`): pinned H2D of every search image,
          on-device Pillow-exact crop/resize pipeline, D2H of the per-crop results, inside the timed region
N > 1: searches are sharded over ranks (independent units, no data-path collective; weights replicated); the time
is the max over ranks of CUDA-event time, value = crops of all ranks / that time ("weak" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CROP = 5.00e12      # BASELINE.md §3 (T=320, g=6, KV-cached); roofline 289 crops/s/GPU at 1443 TF/s


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--searches", type=int, default=32, help="concurrent searches per step per GPU")
    ap.add_argument("--image", type=int, default=1024)
    ap.add_argument("--smallest", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64, help="frontier batch (crops per engine call)")
    ap.add_argument("--tiny", action="store_true", help="tiny model (debug only; not a valid bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attn-impl", type=int, default=0, help="A/B switch for vsb_attn_set_impl (0 = production dispatch)")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed resident steps (for ncu)")
    ap.add_argument("--no-prefix-cache", action="store_true",
                    help="recompute the K/V rows of the constant text prefix (system prompt before <im_start>) for every crop")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["bf16_tflops_sustained"], j["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path on the host cores (bounded sample)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(llama_layers=2, tiny=False):
    """One crop of BASELINE.json configs[0] (768x768 synthetic image, 1-crop VSM forward = the reference's
    VSMForCausalLM.model_forward(inference=True), fp32, random-init weights) through oracle/vsm_oracle.py.
    Bounded: every stage runs in full at full width EXCEPT the 32 identical Llama layers, of which `llama_layers`
    are executed and the measured per-layer time is scaled to 32.  Returns (seconds per crop, description)."""
    import torch
    import torch.nn.functional as F
    from oracle import vsm_oracle as O
    from vstar_b200.config import VSMConfig, tiny_config
    from vstar_b200 import synth
    # all host threads (torchrun exports OMP_NUM_THREADS=1, which would silently make this a 1-core baseline)
    try:
        n_threads = len(os.sched_getaffinity(0))
    except Exception:
        n_threads = os.cpu_count() or 1
    # measured on the pool's boxes: 64 threads -> 12 s per crop, 128 SMT threads -> 180 s (oversubscribed fp32 GEMMs); the
    # baseline uses the faster setting and reports the thread count it used
    n_threads = min(max(1, n_threads), 64)
    torch.set_num_threads(n_threads)
    cfg = tiny_config() if tiny else VSMConfig()
    full_layers = cfg.n_layers
    run_layers = min(llama_layers, full_layers)
    cfg_run = VSMConfig(**{**cfg.to_dict(), "n_layers": run_layers})
    shapes = synth.state_dict_shapes(cfg_run)
    sd = {}
    for name, shape in shapes.items():
        if name == "model.embed_tokens.weight":
            continue                      # the sample feeds embeddings directly (a table lookup costs nothing)
        sd[name] = synth.synthetic_tensor(name, shape, seed=1234)
    T = 320
    g = torch.Generator().manual_seed(0)
    images_clip = torch.randn(1, 3, cfg.clip_image, cfg.clip_image, generator=g)
    images = torch.randn(1, 3, cfg.owl_image, cfg.owl_image, generator=g)
    t = {}

    def timed(key, fn):
        t0 = time.perf_counter()
        r = fn()
        t[key] = time.perf_counter() - t0
        return r

    with torch.no_grad():
        feats = timed("clip+projector", lambda: O.encode_images(sd, cfg_run, images_clip))
        embeds = torch.cat([torch.randn(1, T - feats.shape[1], cfg.hidden, generator=g) * 0.5, feats], dim=1)
        hidden = timed("llama_layers", lambda: O.llama_forward(sd, cfg_run, embeds))
        timed("lm_head_all_rows", lambda: F.linear(hidden, sd["lm_head.weight"]))
        seg = timed("text_fcs_all_rows", lambda: (O.text_fcs(sd, "seg", hidden), O.text_fcs(sd, "det", hidden)))
        fmap = timed("owl_vit", lambda: O.owl_visual_embs(sd, cfg_run, images))
        low = timed("sam_decoder", lambda: O.sam_low_res_masks(sd, cfg_run, fmap[0], seg[0][0, -3:-2]))
        timed("owl_heads", lambda: O.owl_heads(sd, cfg_run, fmap[0], seg[1][0, -3:-2]))
        timed("heatmap", lambda: F.interpolate(low.float(), (768, 768), mode="bilinear", align_corners=False).clamp(min=0).max())
    per_layer = t["llama_layers"] / run_layers
    total = sum(v for k, v in t.items() if k != "llama_layers") + per_layer * full_layers
    desc = (f"1 crop of configs[0] (768x768, T={T}, fp32, oracle port of model_forward(inference=True)); all stages in full, "
            f"{run_layers}/{full_layers} Llama layers executed and scaled x{full_layers / run_layers:g}; stage seconds "
            + json.dumps({k: round(v, 3) for k, v in t.items()}))
    return total, desc, torch.get_num_threads()


def torch_gpu_reference_sample(cfg, n_crops=3):
    """The reference's AS-WRITTEN GPU path restated with plain torch eager ops (oracle/vsm_oracle.py on cuda:0, bf16):
    batch 1, `generate(use_cache=False)` = one full CLIP + 7B pass per emitted token (5 here), lm_head and both query MLPs
    on all T rows, OWL-ViT, SAM decoder, heads, heat-map (VSM.py:438-553).  This is the "reference PyTorch-GPU" number the
    north-star's >= 10x is measured against; the reference itself cannot travel to the GPU box (Python sources under
    /root/reference + transformers 4.31), so its restatement stands in.  Host preprocessing excluded (favours the baseline)."""
    import torch
    from oracle import vsm_oracle as O
    from vstar_b200 import synth
    shapes = synth.state_dict_shapes(cfg)
    sd = {n: synth.synthetic_tensor(n, shp, seed=1234, dtype=torch.bfloat16, device="cuda") for n, shp in shapes.items()}
    prompt, ans = synth.synthetic_prompt(cfg, n_text=60, seed=0)
    prompt = prompt.cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    ic = torch.randn(1, 3, cfg.clip_image, cfg.clip_image, device="cuda", generator=g).bfloat16()
    io = torch.randn(1, 3, cfg.owl_image, cfg.owl_image, device="cuda", generator=g).bfloat16()
    with torch.no_grad():
        O.vsm_inference(sd, cfg, ic, io, prompt, (512, 512), max_new_tokens=100, mode="detection", forced_ids=ans)     # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_crops):
            out = O.vsm_inference(sd, cfg, ic, io, prompt, (512, 512), max_new_tokens=100, mode="detection", forced_ids=ans)
            _ = out["pred_masks"].clamp(min=0).max().item()
        e1.record()
        torch.cuda.synchronize()
    del sd
    torch.cuda.empty_cache()
    return n_crops / (e0.elapsed_time(e1) / 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    times, last = [], None
    desc, cores = "", torch.get_num_threads()
    t_wall = time.perf_counter()
    for i in range(args.warmup + args.steps):
        last, desc, cores = cpu_reference_sample(llama_layers=1 if not args.tiny else 2, tiny=args.tiny)
        if i >= args.warmup:
            times.append(last)
        # keep the whole arm within a few minutes on any host: stop once 240 s of wall clock are spent (a slow box then
        # reports fewer timed samples; "steps" says how many)
        if time.perf_counter() - t_wall > 240:
            break
    if not times:
        times = [last]
    sec = sum(times) / max(1, len(times))
    v = 1.0 / sec
    print(json.dumps({
        "impl": "reference", "metric": "crops/s through the VSM forward (guided visual search hot path)", "value": v, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    return {"workload": f"BASELINE.json configs[1]: {args.image}x{args.image} synthetic images, smallest_size {args.smallest} "
                        f"(root + 4 crops, depth 2), {args.searches} lock-step searches per step per GPU, frontier batch {args.batch}, "
                        "bf16, random-init Vicuna-7B/CLIP-L/OWL-B/SAM-decoder weights, T=315+5 tokens, forced answer ids",
            "searches_per_step": args.searches, "frontier_batch": args.batch, "image": args.image,
            "l2": "weights (13.5 GB) and activations exceed the 126 MB L2 every step; no explicit flush",
            "prefix_cache": ("off" if args.no_prefix_cache else
                             "on: K/V of the 37 constant prompt tokens before <im_start> snapshotted at the first (warm-up) prefill and "
                             "shared by all crops; identical results, see DESIGN.md §3 (--no-prefix-cache recomputes them)")}


# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from PIL import Image

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from vstar_b200 import _lib, ops, synth
    from vstar_b200.config import VSMConfig, tiny_config
    from vstar_b200.engine import VSMEngine, VSMWeights
    from vstar_b200.visual_search import visual_search_many
    from vstar_b200.vsm import VSM

    if args.attn_impl:
        _lib.call("vsb_attn_set_impl", args.attn_impl)
    cfg = tiny_config() if args.tiny else VSMConfig()
    t0 = time.time()
    weights = VSMWeights(cfg, lambda n, _s=synth.state_dict_shapes(cfg): synth.synthetic_tensor(n, _s[n], seed=1234, device="cuda"))
    engine = VSMEngine(weights, max_tokens=384)
    engine.prefix_cache = not args.no_prefix_cache
    prompt, ans = synth.synthetic_prompt(cfg, n_text=60, seed=0, im_start_index=None if args.tiny else 37)

    class BenchVSM(VSM):
        """synthetic tokenisation (no sentencepiece model offline): fixed 60-id prompt, forced 5-id answer"""

        def _ids(self, question):
            return prompt[0].tolist()

    vsm = BenchVSM(engine=engine, forced_answer_ids=ans.tolist(), frontier_batch=args.batch)
    load_s = time.time() - t0

    S = args.searches
    images = []
    for i in range(S):
        arr = np.random.default_rng(1000 * rank + i).integers(0, 256, (args.image, args.image, 3), dtype=np.uint8)
        images.append(Image.fromarray(arr, "RGB"))
    jobs = [(im, "mug", args.smallest) for im in images]
    kw = dict(confidence_high=2.0, target_cue_threshold=-1e9, target_cue_threshold_minimum=-1e9)

    def step(resident=False):
        if not resident:
            vsm.release()          # public path: every search image is uploaded (pinned H2D) inside the step
        res, states = visual_search_many(vsm, jobs, batch_size=args.batch, **kw)
        return sum(st.n_evals for st in states), states

    def timed_steps(n, resident=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        crops = 0
        for _ in range(n):
            c, _ = step(resident)
            crops += c
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            c = torch.tensor([crops], device="cuda", dtype=torch.float64)
            dist.all_reduce(c)
            crops = int(c[0])
        return ms, crops

    # ---- e2e leg: public API from PIL images (upload of the search images, GPU crop/resize pipeline, D2H of results)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    h2d0 = vsm.h2d_bytes
    e2e_ms, e2e_crops = timed_steps(args.steps)
    bytes_h2d = (vsm.h2d_bytes - h2d0) // args.steps
    crops_per_step_rank = e2e_crops // args.steps // world
    bytes_d2h = crops_per_step_rank * (5 * 4) + (crops_per_step_rank // 5) * 2 * (3 * 4 + 5 * 8)   # top score+box per crop; stats+rect sums per expansion

    # ---- device-resident leg: same searches, search images already in HBM
    for _ in range(max(1, args.warmup)):
        step(True)
    # GEMM roofline: CUDA events around every tcgen05 GEMM launch of the timed region
    ops.profile_begin()
    launches0 = _lib.launches
    if args.profile_range:           # `ncu --profile-from-start off`: capture exactly the timed resident steps
        torch.cuda.profiler.start()
    dev_ms, dev_crops = timed_steps(args.steps, True)
    if args.profile_range:
        torch.cuda.profiler.stop()
    launches = _lib.launches - launches0
    gemm_flops, gemm_ms, gemm_n = ops.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    prefix_rows = engine._P

    # transparency leg (all ranks: timed_steps synchronises them): the same resident workload with the shared-prefix KV and
    # the tail-only last layer switched OFF, i.e. all 320 rows of every crop through all 32 layers (5.00 TFLOP/crop)
    value_full = None
    if not args.no_prefix_cache:
        saved = (engine.prefix_cache, engine.tail_only)
        try:
            engine.prefix_cache, engine.tail_only = False, False
            step(True)
            ms_full, crops_full = timed_steps(min(2, args.steps), True)
            value_full = crops_full / (ms_full / 1e3)
        except Exception:
            value_full = None
        finally:
            engine.prefix_cache, engine.tail_only = saved

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_hbm, peak_src = peaks()
    # executed FLOPs per crop: BASELINE.md §3's 5.00e12 minus the 7B linear work of the prefix rows that are not recomputed
    flops_per_crop = FLOPS_PER_CROP - 2.0 * prefix_rows * 6.476e9
    if engine.tail_only and not args.tiny:      # last layer: o-proj + MLP only on the 5 consumed rows of each crop
        flops_per_crop -= 2.0 * (320 - prefix_rows - 5) * (cfg.hidden * cfg.hidden + 3 * cfg.hidden * cfg.intermediate)
    value = dev_crops / (dev_ms / 1e3)
    e2e = e2e_crops / (e2e_ms / 1e3)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    out = {
        "metric": "crops/s through the VSM forward (guided visual search hot path)", "value": value, "unit": "crops/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e, "unit": "crops/s", "h2d_bytes_per_step": bytes_h2d, "d2h_bytes_per_step": bytes_d2h,
                "ms_per_step": e2e_ms / args.steps, "host_prep_s_total": vsm.timers["prep"]},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": None, "kernel": "gemm_bf16_tcgen05_kernel", "launches": gemm_n,
                     "how": "sum(2*M*N*K) / sum(CUDA-event duration) over every GEMM launch of the timed region; peak = " + peak_src,
                     "whole_path_frac": value / world * flops_per_crop / (peak_tf * 1e12),
                     "whole_path_note": ("crops/s/GPU x %.2f TFLOP/crop executed / peak (BASELINE.md §3 counts 5.00 TFLOP/crop at T=320; "
                                         "%d constant prefix rows per crop are served from the shared-prefix KV snapshot, and the last "
                                         "decoder layer runs o-proj/MLP on the 5 consumed rows only)"
                                         % (flops_per_crop / 1e12, prefix_rows))},
        "value_all_rows_all_layers": value_full,      # same leg with prefix sharing and the tail-only last layer off (5.00 TFLOP/crop)
        "crops_per_step": dev_crops // args.steps, "load_s": load_s,
        "draft_verify": engine.stats,
    }
    out["roofline"]["traffic"] = 1.528e9
    out["roofline"]["traffic_note"] = ("dram__bytes_read+write of the dominant launch (gate|up GEMM of the 32-crop batch, M=10240 N=22016 K=4096; "
                                       "algorithmic 0.489e9 B) from profiles/r01_gemm_2cta_gateup_M10240_ncu_details.csv")
    if not args.no_cpu_baseline and world == 1:
        try:
            # latency of ONE visual_search() call on its own (root + 4 crops), public API
            vsm.release()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            visual_search_many(vsm, jobs[:1], batch_size=args.batch, **kw)
            torch.cuda.synchronize()
            out["search_latency_ms"] = (time.perf_counter() - t0) * 1e3
            out["torch_gpu_baseline"] = {"value": torch_gpu_reference_sample(cfg), "unit": "crops/s",
                                         "what": "reference as-written GPU path restated in plain torch eager bf16 (batch 1, uncached greedy "
                                                 "generate = 5 full CLIP+7B passes, lm_head/fcs on all rows), cuda:0, host prep excluded"}
        except Exception as e:
            out["torch_gpu_baseline"] = {"value": None, "what": f"failed: {e!r}"}
        try:
            sec, desc, cores = cpu_reference_sample(llama_layers=1 if not args.tiny else 2, tiny=args.tiny)
            out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "crops/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as e:   # the baseline leg must never take the GPU number down
            out["cpu_baseline"] = {"value": None, "unit": "crops/s", "cores": None, "kind": "port", "sample": f"failed: {e!r}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
