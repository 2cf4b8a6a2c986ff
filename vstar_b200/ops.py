"""Tensor-level wrappers over the C-ABI (torch is only the owner of device memory and streams here).

Every function launches hand-written sm_100a kernels from libvstar_b200.so on the current torch CUDA
stream.  No function has a PyTorch fallback: on a box without the built library, or on CPU tensors,
they raise.
"""
import torch

from . import _lib
from ._lib import call

EPI_NONE, EPI_QUICK_GELU, EPI_GELU, EPI_RELU, EPI_SWIGLU = 0, 1, 2, 3, 4
BF16 = torch.bfloat16


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype=None):
    if not t.is_cuda:
        raise _lib.VsbError("vstar_b200 ops need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.VsbError(f"expected {dtype}, got {t.dtype}")
    return t


def _rows2d(t):
    """(ptr, rows, cols, ld) of a 2-D row-major view with unit inner stride."""
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)


class batch_invariant:
    """with ops.batch_invariant(): kernels are chosen independently of the row count of a call (see vsb_set_batch_invariant)"""
    depth = 0

    def __enter__(self):
        if batch_invariant.depth == 0:
            call("vsb_set_batch_invariant", 1)
        batch_invariant.depth += 1

    def __exit__(self, *a):
        batch_invariant.depth -= 1
        if batch_invariant.depth == 0:
            call("vsb_set_batch_invariant", 0)


# optional per-launch profiling of the dominant kernel (bench.py roofline): CUDA events around every GEMM launch
profiling = False


def profile_begin():
    """CUDA events around every vsb_gemm_bf16 launch from here on (recorded inside the library, so launches issued by the native
    layer runner are covered too)"""
    global profiling
    profiling = True
    call("vsb_gemm_profile_begin")


def profile_end():
    """-> (total algorithmic flops, total ms, launches) over the GEMM launches since profile_begin()"""
    import ctypes
    global profiling
    f, t, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    call("vsb_gemm_profile_end", ctypes.byref(f), ctypes.byref(t), ctypes.byref(n))
    profiling = False
    return f.value, t.value, n.value


def llama_layer_table(layers):
    """layers: list of dicts with ln1, wqkv, wo, ln2, wgu, wdown tensors (kept alive by the caller) -> ctypes array"""
    arr = (_lib.LlamaLayer * len(layers))()
    for i, L in enumerate(layers):
        for k in ("ln1", "wqkv", "wo", "ln2", "wgu", "wdown"):
            t = L[k]
            assert t.dtype == BF16 and t.is_contiguous() and t.is_cuda, k
            setattr(arr[i], k, t.data_ptr())
    return arr


def llama_layers(table, n_layers, x, B, Tn, past, cache, Bc, Tmax, d, H, inter, eps, rope_cos, rope_sin, scratch, positions=None,
                 k_start=None, cache_row_offset=0, tail_rows=0, q_seg=None, seg_lo=0, norm_folded=False):
    """whole decoder stack in ONE native call (csrc/llama_layers.cu): in place on x [B*Tn, d].
    cache_row_offset shifts the cache origin by that many [3d] rows (a sequence placed at batch slot b, first row s:
    b*Tmax + s) so a single sequence can be prefilled anywhere in a shared cache; positions / k_start: ragged decode;
    tail_rows > 0: only the last tail_rows rows of every sequence are valid on return (the last layer skips the others);
    q_seg int32 [B*Tn] + seg_lo: the new rows are several continuations of the cached prefix [0, seg_lo) appended back to back;
    row r attends the prefix and its own continuation from key q_seg[r] on;
    norm_folded: ln1 / ln2 are already multiplied into wqkv / wgu (fused RMSNorm path, see vsb_llama_layers)."""
    _chk(x, BF16), _chk(cache, BF16), _chk(scratch, BF16)
    assert x.is_contiguous() and cache.is_contiguous() and x.shape == (B * Tn, d)
    assert scratch.numel() >= B * Tn * (2 * d + inter)
    for t in (positions, k_start, q_seg):
        assert t is None or (t.dtype == torch.int32 and t.is_contiguous() and t.is_cuda)
    assert q_seg is None or (q_seg.numel() == B * Tn and positions is not None)
    assert positions is None or positions.numel() == B * Tn
    assert k_start is None or k_start.numel() == B
    _lib.launches += (6 if (norm_folded and k_start is None and (B * Tn > 16 or batch_invariant.depth > 0)) else 8) * n_layers + \
        (2 if (tail_rows > 0 and 2 * tail_rows <= Tn and k_start is None) else 0)
    call("vsb_llama_layers", table, n_layers, x.data_ptr(), B, Tn, past, cache.data_ptr() + cache_row_offset * 3 * d * 2, Bc, Tmax, d, H,
         inter, float(eps), rope_cos.data_ptr(), rope_sin.data_ptr(), _p(positions), _p(k_start), int(tail_rows), _p(q_seg), int(seg_lo),
         1 if norm_folded else 0, scratch.data_ptr(), _stream())
    return x


def gemm(a, w, out=None, bias=None, residual=None, epilogue=EPI_NONE, out_dtype=BF16, rows_per_group=0,
         group_stride=0, group_offset=0, out_rows=None):
    """out = epi(a @ w.T + bias) (+ residual).  a [M,K], w [N,K] bf16 (row stride arbitrary, inner stride 1)."""
    _chk(a, BF16), _chk(w, BF16)
    pa, M, K, lda = _rows2d(a)
    pw, N, K2, ldw = _rows2d(w)
    assert K == K2, (a.shape, w.shape)
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, n_out), dtype=out_dtype, device=a.device)
    _chk(out)
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[1] >= n_out
    out_fp32 = 1 if out.dtype == torch.float32 else 0
    if residual is not None:
        assert residual.dtype == out.dtype and residual.stride(1) == 1
    if bias is not None:
        _chk(bias, BF16)
        assert bias.numel() == N and bias.is_contiguous()
    _lib.launches += 1
    call("vsb_gemm_bf16", pa, lda, pw, ldw, out.data_ptr(), out.stride(0), M, N, K, _p(bias), _p(residual),
         residual.stride(0) if residual is not None else 0, epilogue, out_fp32, rows_per_group, group_stride, group_offset,
         _stream())
    return out


def gemm_rowscale(a, w, out, rowsq_in=None, rowsq_out=None, eps=1e-6, residual=None, epilogue=EPI_NONE):
    """out = epi(rstd[m] * (a @ w.T)) (+ residual) with the RMSNorm weight already folded into w; rowsq_in fp32 [chunks, M] partial
    sums of squares of a's rows, rowsq_out fp32 [N/32, M] receives those of the stored output (see vsb_gemm_rowscale_bf16)"""
    _chk(a, BF16), _chk(w, BF16), _chk(out, BF16)
    pa, M, K, lda = _rows2d(a)
    pw, N, K2, ldw = _rows2d(w)
    assert K == K2 and out.stride(1) == 1
    for t in (rowsq_in, rowsq_out):
        assert t is None or (t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.shape[1] == M)
    _lib.launches += 1
    call("vsb_gemm_rowscale_bf16", pa, lda, pw, ldw, out.data_ptr(), out.stride(0), M, N, K, 0, _p(residual),
         residual.stride(0) if residual is not None else 0, epilogue, 0, 0, 0, _p(rowsq_in), rowsq_in.shape[0] if rowsq_in is not None else 0,
         float(eps), _p(rowsq_out), M, _stream())
    return out


def rowsq(x):
    """fp32 [1, rows] sums of squares of the rows of x (bf16)"""
    _chk(x, BF16)
    px, rows, cols, ldx = _rows2d(x)
    out = torch.empty((1, rows), dtype=torch.float32, device=x.device)
    _lib.launches += 1
    call("vsb_rowsq_bf16", px, ldx, out.data_ptr(), rows, cols, _stream())
    return out


def layernorm(x, w, b, eps, out=None, act=EPI_NONE):
    _chk(x, BF16)
    px, rows, cols, ldx = _rows2d(x)
    if out is None:
        out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_layernorm_bf16", px, ldx, w.data_ptr(), b.data_ptr(), out.data_ptr(), out.stride(0), rows, cols, float(eps), act, _stream())
    return out


def rmsnorm(x, w, eps, out=None):
    _chk(x, BF16)
    px, rows, cols, ldx = _rows2d(x)
    if out is None:
        out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_rmsnorm_bf16", px, ldx, w.data_ptr(), out.data_ptr(), out.stride(0), rows, cols, float(eps), _stream())
    return out


def rope_(qkv, T, H, D, cos_t, sin_t, pos0=0, positions=None, rows=None, group_stride=None, group_offset=0):
    """In place on the q|k thirds of qkv [.., 3*H*D].  Default: `rows` = all rows, contiguous groups of T.
    KV-cache layout: logical row r -> physical row (r // T) * group_stride + group_offset + r % T."""
    _chk(qkv, BF16)
    p, nrows, cols, ld = _rows2d(qkv)
    assert cols == 3 * H * D
    if rows is None:
        rows = nrows
    if group_stride is None:
        group_stride = T
    _lib.launches += 1
    call("vsb_rope_bf16", p, ld, rows, T, H, D, pos0, cos_t.data_ptr(), sin_t.data_ptr(), _p(positions), group_stride, group_offset,
         _stream())
    return qkv


def embed_splice(ids, table, out, img_pos, n_img):
    B, L = ids.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous() and out.is_contiguous()
    _lib.launches += 1
    call("vsb_embed_splice_bf16", ids.data_ptr(), table.data_ptr(), out.data_ptr(), B, L, img_pos, n_img, table.shape[1], table.shape[0], _stream())
    return out


def gather_rows(idx, table, out=None):
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    pt, nrows, d, ldt = _rows2d(table)
    n = idx.numel()
    if out is None:
        out = torch.empty((n, d), dtype=BF16, device=table.device)
    _lib.launches += 1
    call("vsb_gather_rows_bf16", idx.data_ptr(), pt, ldt, out.data_ptr(), out.stride(0), n, d, nrows, _stream())
    return out


def patchify(pixels, P, Kpad):
    """pixels [B,3,S,S] bf16 -> [B*g*g, Kpad]"""
    _chk(pixels, BF16)
    B, C, S, S2 = pixels.shape
    assert C == 3 and S == S2 and pixels.is_contiguous()
    g = S // P
    out = torch.empty((B * g * g, Kpad), dtype=BF16, device=pixels.device)
    _lib.launches += 1
    call("vsb_patchify_bf16", pixels.data_ptr(), out.data_ptr(), B, S, P, Kpad, _stream())
    return out


def vit_add_pos_(x, cls, pos, B, S):
    assert x.is_contiguous()
    _lib.launches += 1
    call("vsb_vit_add_pos_bf16", x.data_ptr(), cls.data_ptr(), pos.data_ptr(), B, S, x.shape[-1], _stream())
    return x


def owl_merge(x, w1, b1, w2, b2, B, S, eps):
    C = x.shape[-1]
    assert x.is_contiguous()
    out = torch.empty((B * (S - 1), C), dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_owl_merge_bf16", x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out.data_ptr(), B, S, C, float(eps), _stream())
    return out


def add_rows(a, b, out=None):
    """out[r] = a[r] + b[r % b.rows] (bf16)."""
    assert a.is_contiguous() and b.is_contiguous() and a.shape[-1] == b.shape[-1]
    a2, b2 = a.view(-1, a.shape[-1]), b.view(-1, b.shape[-1])
    if out is None:
        out = torch.empty_like(a)
    _lib.launches += 1
    call("vsb_add_rows_bf16", a2.data_ptr(), b2.data_ptr(), out.data_ptr(), a2.shape[0], a2.shape[1], b2.shape[0], _stream())
    return out


def cast_f32_bf16(x):
    _chk(x, torch.float32)
    assert x.is_contiguous()
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_cast_f32_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def argmax_rows(x):
    _chk(x, torch.float32)
    p, rows, n, ld = _rows2d(x)
    idx = torch.empty((rows,), dtype=torch.int32, device=x.device)
    val = torch.empty((rows,), dtype=torch.float32, device=x.device)
    _lib.launches += 1
    call("vsb_argmax_rows_f32", p, ld, rows, n, idx.data_ptr(), val.data_ptr(), _stream())
    return idx, val


def nll_rows(logits, labels):
    """logits fp32 [rows, V], labels int64 [rows] -> nll fp32 [rows]"""
    _chk(logits, torch.float32)
    p, rows, n, ld = _rows2d(logits)
    assert labels.dtype == torch.int64 and labels.numel() == rows and labels.is_contiguous()
    out = torch.empty((rows,), dtype=torch.float32, device=logits.device)
    _lib.launches += 1
    call("vsb_nll_rows_f32", p, ld, rows, n, labels.data_ptr(), out.data_ptr(), _stream())
    return out


def copy2d(src, dst):
    """dst[:, :] = src[:, :] for 2-byte element 2-D views with unit inner stride."""
    ps, rows, cols, lds = _rows2d(src)
    pd, rows2, cols2, ldd = _rows2d(dst)
    assert rows == rows2 and cols == cols2 and src.element_size() == 2 and dst.element_size() == 2
    _lib.launches += 1
    call("vsb_copy2d_b16", ps, lds, pd, ldd, rows, cols, _stream())
    return dst


def flash_attn(q, k, v, out, B, H, Sq, Sk, D, causal, scale, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs):
    """Strided attention: element (b,s,h,d) of q at q.data_ptr() + (b*q_bs + s*q_rs + h*D + d) elements, etc."""
    _lib.launches += 1
    call("vsb_flash_attn_bf16", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), q_bs, q_rs, k_bs, k_rs, v_bs, v_rs,
         o_bs, o_rs, B, H, Sq, Sk, D, 1 if causal else 0, float(scale), _stream())
    return out


def attn_fused_qkv(qkv, B, S, H, D, causal, scale, out=None):
    """qkv [B*S, 3*H*D] (q|k|v) -> out [B*S, H*D]"""
    ld = qkv.stride(0)
    if out is None:
        out = torch.empty((B * S, H * D), dtype=BF16, device=qkv.device)
    q = qkv
    k = qkv[:, H * D:]
    v = qkv[:, 2 * H * D:]
    return flash_attn(q, k, v, out, B, H, S, S, D, causal, scale, S * ld, ld, S * ld, ld, S * ld, ld, S * out.stride(0), out.stride(0))


def attn_small(q, k, v, B, H, Nq, Nk, D, scale, out=None):
    pq, rq, cq, ldq = _rows2d(q)
    pk, rk, ck, ldk = _rows2d(k)
    pv, rv, cv, ldv = _rows2d(v)
    if out is None:
        out = torch.empty((B * Nq, H * D), dtype=BF16, device=q.device)
    _lib.launches += 1
    call("vsb_attn_small_bf16", pq, ldq, pk, ldk, pv, ldv, out.data_ptr(), out.stride(0), B, H, Nq, Nk, D, float(scale), _stream())
    return out


def owl_class_post(y, query, rows_per_crop, Q, quant_bf16=False):
    """quant_bf16: round logits / scores to bf16 values (what the reference's bf16 model emits), fp32 storage"""
    _chk(y, torch.float32), _chk(query, BF16)
    R = y.shape[0]
    logits = torch.empty((R,), dtype=torch.float32, device=y.device)
    scores = torch.empty((R,), dtype=torch.float32, device=y.device)
    _lib.launches += 1
    call("vsb_owl_class_post", y.data_ptr(), y.stride(0), query.data_ptr(), query.stride(0), rows_per_crop, R, Q, 1 if quant_bf16 else 0,
         logits.data_ptr(), scores.data_ptr(), _stream())
    return logits, scores


def owl_box_post(y, box_bias, rows_per_crop, quant_bf16=False):
    _chk(y, torch.float32), _chk(box_bias, torch.float32)
    R = y.shape[0]
    boxes = torch.empty((R, 4), dtype=torch.float32, device=y.device)
    _lib.launches += 1
    call("vsb_owl_box_post", y.data_ptr(), y.stride(0), box_bias.data_ptr(), rows_per_crop, R, 1 if quant_bf16 else 0, boxes.data_ptr(),
         _stream())
    return boxes


def pack_detections(scores, boxes, rec, row0=0):
    """scores fp32 [n,P], boxes fp32 [n,P,4] -> records rec[row0 : row0+n] (fields 0..7 and 12..75, see records.py)"""
    _chk(scores, torch.float32), _chk(boxes, torch.float32), _chk(rec, torch.float32)
    n, P = scores.shape
    assert scores.is_contiguous() and boxes.is_contiguous() and boxes.shape == (n, P, 4)
    assert rec.dim() == 2 and rec.is_contiguous() and row0 + n <= rec.shape[0]
    _lib.launches += 1
    call("vsb_pack_detections_f32", scores.data_ptr(), boxes.data_ptr(), n, P, rec.data_ptr() + row0 * rec.shape[1] * 4, rec.shape[1],
         _stream())
    return rec


def heat_pyramids(jobs, rec, LH, LW):
    """jobs: list of (low_res fp32 [LH,LW] device view, h, w, rects [(x,y,w,h) relative to the crop], record row).
    Fills the heat-map statistics and rectangle sums of those records straight from the low-res masks (no H x W map)."""
    import numpy as np
    if not jobs:
        return rec
    _chk(rec, torch.float32)
    R = rec.shape[1]
    nj = len(jobs)
    total = sum(len(j[3]) for j in jobs)
    table = np.zeros(nj * 8 + total * 5, dtype=np.int32)
    jt = table[:nj * 8].reshape(nj, 8)
    rt = table[nj * 8:nj * 8 + total * 4].reshape(total, 4)
    rj = table[nj * 8 + total * 4:]
    off = 0
    for k, (low, h, w, rects, row) in enumerate(jobs):
        assert low.dtype == torch.float32 and low.is_cuda and low.is_contiguous() and low.shape[-2:] == (LH, LW)
        assert len(rects) <= R - 76, (len(rects), R)
        p = low.data_ptr()
        jt[k] = (0, 0, h, w, off, len(rects), row, 0)
        jt[k, 0:2] = np.array([p], dtype=np.uint64).view(np.int32)
        if rects:
            rt[off:off + len(rects)] = np.asarray(rects, dtype=np.int32)
            rj[off:off + len(rects)] = k
        off += len(rects)
    dev = rec.device
    tab = torch.from_numpy(table).to(dev, non_blocking=True)
    stats_scratch = torch.empty((nj * 64 * 3,), dtype=torch.float32, device=dev)
    rect_scratch = torch.empty((max(1, total) * 64,), dtype=torch.float64, device=dev)
    base = tab.data_ptr()
    _lib.launches += 2 + (2 if total else 0)
    call("vsb_heat_pyramids_f32", base, nj, base + nj * 32, base + nj * 32 + total * 16, total, LH, LW, rec.data_ptr(), R,
         stats_scratch.data_ptr(), rect_scratch.data_ptr(), _stream())
    return rec


def upsample2x_nhwc(x, B, H, W, C):
    assert x.is_contiguous()
    out = torch.empty((B * 4 * H * W, C), dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_upsample2x_nhwc_bf16", x.data_ptr(), out.data_ptr(), B, H, W, C, _stream())
    return out


def im2col3x3_nhwc(x, B, H, W, C):
    assert x.is_contiguous()
    out = torch.empty((B * H * W, 9 * C), dtype=BF16, device=x.device)
    _lib.launches += 1
    call("vsb_im2col3x3_nhwc_bf16", x.data_ptr(), out.data_ptr(), B, H, W, C, _stream())
    return out


def mask_dot(up, hyper, B, P, C):
    assert up.is_contiguous() and hyper.is_contiguous()
    out = torch.empty((B, P), dtype=torch.float32, device=up.device)
    _lib.launches += 1
    call("vsb_mask_dot_bf16", up.data_ptr(), hyper.data_ptr(), out.data_ptr(), B, P, C, _stream())
    return out


HEATMAP_MAX_BLOCKS = 1184


def heatmap(low, h, w, clamp=True, with_stats=True, out=None):
    """low [LH,LW] fp32 -> (map [h,w] fp32, stats [3] = max,min,sum)"""
    _chk(low, torch.float32)
    assert low.is_contiguous() and low.dim() == 2
    if out is None:
        out = torch.empty((h, w), dtype=torch.float32, device=low.device)
    stats = scratch = None
    if with_stats:
        stats = torch.empty((3,), dtype=torch.float32, device=low.device)
        scratch = torch.empty((3 * HEATMAP_MAX_BLOCKS,), dtype=torch.float32, device=low.device)
    _lib.launches += 2 if with_stats else 1
    call("vsb_heatmap_bilinear_f32", low.data_ptr(), low.shape[0], low.shape[1], out.data_ptr(), h, w, 1 if clamp else 0,
         _p(scratch), _p(stats), _stream())
    return out, stats


def rect_sums(hm, rects, stats):
    """hm [h,w] fp32, rects int32 [n,4] (x,y,w,h), stats [3] -> float64 [n] sums of the normalised map"""
    assert rects.dtype == torch.int32 and rects.is_contiguous()
    n = rects.shape[0]
    out = torch.empty((n,), dtype=torch.float64, device=hm.device)
    scratch = torch.empty((64 * n,), dtype=torch.float64, device=hm.device)
    _lib.launches += 2
    call("vsb_rect_sums_f32", hm.data_ptr(), hm.shape[0], hm.shape[1], rects.data_ptr(), n, stats.data_ptr(), out.data_ptr(),
         scratch.data_ptr(), _stream())
    return out
