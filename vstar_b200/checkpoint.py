"""Checkpoint reader for the reference's wire format: HF sharded `*.safetensors` / `pytorch_model-*.bin` with
the key layout written by /root/reference/VisualSearch/merge_lora_weights_and_save_hf_model.py:143-151 (`model.layers.*`,
`model.mm_projector.*`, `model.owlvit.*`, `model.visual_projection.*`, `model.prompt_encoder.*`, `model.mask_decoder.*`,
`model.text_hidden_fcs_{seg,det}.*`, `lm_head.*`; the CLIP tower is a separate checkpoint) and `seal_vqa_7b`
(`model.mm_projector`, `model.mm_projector_object.*`).

safetensors shards are STREAMED to the GPU: the file is memory-mapped, each tensor's bytes go through a small ring of pinned
staging buffers and `cudaMemcpyAsync` on a copy stream, so the host's page-cache / disk read of chunk k+1 overlaps the DMA of
chunk k and no shard is ever materialised in host RAM (the reference's `from_pretrained(low_cpu_mem_usage=True)` still builds
every tensor on the host first, visual_search.py:157-159).  `.bin` shards (pickles) are opened with `mmap=True`, one shard at a
time, located through `pytorch_model.bin.index.json` when it exists.
"""
from __future__ import annotations

import glob
import json
import os
import struct
import time

import numpy as np
import torch

_ST_DTYPES = {"F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16, "I64": torch.int64,
              "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}


class SafetensorsStream:
    """name -> tensor on `device`, streamed from memory-mapped safetensors shards through pinned staging buffers"""

    def __init__(self, files, device="cuda", chunk_bytes=32 << 20, slots=3):
        self.device = torch.device(device)
        self.chunk = int(chunk_bytes)
        self.index = {}                       # name -> (file id, dtype, shape, begin, end)
        self.maps = []
        for fi, path in enumerate(files):
            with open(path, "rb") as f:
                (n,) = struct.unpack("<Q", f.read(8))
                header = json.loads(f.read(n))
            base = 8 + n
            self.maps.append(np.memmap(path, dtype=np.uint8, mode="r"))
            for name, meta in header.items():
                if name == "__metadata__":
                    continue
                b, e = meta["data_offsets"]
                self.index[name] = (fi, _ST_DTYPES[meta["dtype"]], tuple(meta["shape"]), base + b, base + e)
        self.stats = dict(tensors=0, bytes=0, seconds=0.0)
        self._slots = [[None, None] for _ in range(slots)]          # [pinned uint8 tensor, event of the last copy out of it]
        self._next = 0
        self._stream = None

    def keys(self):
        return self.index.keys()

    def __contains__(self, name):
        return name in self.index

    def __call__(self, name):
        return self.get(name)

    def get(self, name):
        fi, dtype, shape, b, e = self.index[name]
        src = self.maps[fi]
        t0 = time.perf_counter()
        if self.device.type != "cuda":
            out = torch.frombuffer(bytearray(src[b:e].tobytes()), dtype=dtype).reshape(shape) if e > b else torch.empty(shape, dtype=dtype)
            self._account(e - b, t0)
            return out
        out = torch.empty(shape, dtype=dtype, device=self.device)
        if e > b:
            flat = out.view(-1).view(torch.uint8)
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream(self.device)
            self._stream.wait_stream(main)               # `out` was allocated on the compute stream
            last = None
            for off in range(b, e, self.chunk):
                n = min(self.chunk, e - off)
                slot = self._slots[self._next % len(self._slots)]
                self._next += 1
                if slot[1] is not None:
                    slot[1].synchronize()                # the DMA that last read this staging buffer has finished
                if slot[0] is None:
                    slot[0] = torch.empty(self.chunk, dtype=torch.uint8).pin_memory()
                np.copyto(slot[0].numpy()[:n], src[off:off + n])          # page-cache / disk read, overlaps the previous chunk's DMA
                with torch.cuda.stream(self._stream):
                    flat[off - b:off - b + n].copy_(slot[0][:n], non_blocking=True)
                    last = torch.cuda.Event()
                    last.record(self._stream)
                slot[1] = last
            out.record_stream(self._stream)
            main.wait_event(last)                        # kernels that consume the tensor are ordered after its last chunk
        self._account(e - b, t0)
        return out

    def _account(self, nbytes, t0):
        self.stats["tensors"] += 1
        self.stats["bytes"] += nbytes
        self.stats["seconds"] += time.perf_counter() - t0


class BinShards:
    """name -> tensor from `pytorch_model*.bin` pickles, one memory-mapped shard open at a time"""

    def __init__(self, files, path):
        self.files = files
        self.weight_map = None
        idx = os.path.join(path, "pytorch_model.bin.index.json")
        if os.path.exists(idx):
            self.weight_map = {k: os.path.join(path, v) for k, v in json.load(open(idx))["weight_map"].items()}
        self._open = (None, None)

    def _load(self, f):
        if self._open[0] != f:
            self._open = (None, None)
            try:
                sd = torch.load(f, map_location="cpu", weights_only=True, mmap=True)
            except Exception:                      # legacy (non-zipfile) pickles cannot be memory-mapped
                sd = torch.load(f, map_location="cpu", weights_only=True)
            self._open = (f, sd)
        return self._open[1]

    def __call__(self, name):
        if self.weight_map is not None:
            return self._load(self.weight_map[name])[name]
        order = ([self._open[0]] if self._open[0] else []) + [f for f in self.files if f != self._open[0]]
        for f in order:
            sd = self._load(f)
            if name in sd:
                return sd[name]
        raise KeyError(name)


def open_checkpoint(path, device="cpu"):
    """-> callable name -> tensor, over *.safetensors or pytorch_model*.bin shards in `path`.  device="cuda": safetensors shards
    are streamed straight to the GPU (see module docstring); .bin tensors come back as host (mmap) tensors."""
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if files:
        return SafetensorsStream(files, device=device)
    files = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {path}")
    return BinShards(files, path)
