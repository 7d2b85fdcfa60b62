"""torch.library registration of the C-ABI kernels as ``torch.ops.wm_b200.*`` (CUDA dispatch key only).

PyTorch here is plumbing: it owns device memory and streams; every op below passes raw device pointers to
libwm_b200.so (ctypes) on ``torch.cuda.current_stream()``.  All ops write into caller-provided tensors
(kernels never allocate), so they are declared as mutating ops returning ``()``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as _lib

_LIBRARY = torch.library.Library("wm_b200", "DEF")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Optional[torch.Tensor], dtype, name: str, inner_contig: bool = True) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.WmError(f"{name}: expected a CUDA tensor (wildlifemapper_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.WmError(f"{name}: expected {dtype}, got {t.dtype}")
    if inner_contig and t.dim() > 0 and t.numel() > 0 and t.stride(-1) != 1:
        raise _lib.WmError(f"{name}: innermost dimension must be contiguous")


def _on_tensor_device(fn):
    """Launch on the device (and that device's current stream) of the op's tensors, not on whatever device happens to be
    current: the C ABI works on the caller's current device."""
    def run(*args):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args)
                break
        return fn(*args)
    return run


def _define(schema: str, fn) -> None:
    name = schema.split("(", 1)[0]
    _LIBRARY.define(schema)
    _LIBRARY.impl(name, _on_tensor_device(fn), "CUDA")


# ------------------------------------------------------------------ GEMM family
def _gemm(a, w, bias, residual, res_mod, out_bf16, out_f32, act, bn_hint):
    _chk(a, torch.bfloat16, "gemm.a"); _chk(w, torch.bfloat16, "gemm.w")
    _chk(bias, torch.float32, "gemm.bias"); _chk(residual, torch.float32, "gemm.residual")
    _chk(out_bf16, torch.bfloat16, "gemm.out_bf16"); _chk(out_f32, torch.float32, "gemm.out_f32")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    for o in (out_bf16, out_f32):
        assert o is None or tuple(o.shape) == (M, N), (o.shape, M, N)
    assert bias is None or bias.numel() == N
    _lib.call("wm_gemm_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), _ptr(residual),
              0 if residual is None else residual.stride(0), int(res_mod), _ptr(out_bf16),
              0 if out_bf16 is None else out_bf16.stride(0), _ptr(out_f32), 0 if out_f32 is None else out_f32.stride(0),
              M, N, K, int(act), int(bn_hint), _stream())


_define("gemm(Tensor a, Tensor w, Tensor? bias, Tensor? residual, int res_mod, Tensor(a!)? out_bf16, "
        "Tensor(b!)? out_f32, int act, int bn_hint) -> ()", _gemm)


def _conv3x3(x, w, out_bf16, out_f32):
    _chk(x, torch.bfloat16, "conv3x3.x"); _chk(w, torch.bfloat16, "conv3x3.w")
    assert x.is_contiguous() and w.is_contiguous()
    B, H, W_, Cc = x.shape
    assert H == 64 and W_ == 64
    N = w.shape[0]
    assert w.shape[1] == 9 * Cc
    for o in (out_bf16, out_f32):
        assert o is None or (o.is_contiguous() and o.numel() == B * 4096 * N)
    _lib.call("wm_conv3x3_nhwc_bf16", x.data_ptr(), w.data_ptr(), _ptr(out_bf16), _ptr(out_f32), B, Cc, N, _stream())


_define("conv3x3(Tensor x, Tensor w, Tensor(a!)? out_bf16, Tensor(b!)? out_f32) -> ()", _conv3x3)


# ------------------------------------------------------------------ bandwidth kernels
def _layernorm(x, gamma, beta, y_bf16, y_f32, add, add_mod, y2_bf16, eps):
    _chk(x, torch.float32, "layernorm.x")
    assert x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    for t in (y_bf16, y_f32, y2_bf16):
        assert t is None or (t.is_contiguous() and t.numel() == x.numel())
    _chk(y_bf16, torch.bfloat16, "layernorm.y_bf16"); _chk(y_f32, torch.float32, "layernorm.y_f32")
    _chk(y2_bf16, torch.bfloat16, "layernorm.y2_bf16"); _chk(add, torch.float32, "layernorm.add")
    _chk(gamma, torch.float32, "layernorm.gamma"); _chk(beta, torch.float32, "layernorm.beta")
    _lib.call("wm_layernorm", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(y_bf16), _ptr(y_f32), _ptr(add),
              int(add_mod), _ptr(y2_bf16), rows, D, float(eps), _stream())


_define("layernorm(Tensor x, Tensor gamma, Tensor beta, Tensor(a!)? y_bf16, Tensor(b!)? y_f32, Tensor? add, "
        "int add_mod, Tensor(c!)? y2_bf16, float eps) -> ()", _layernorm)


def _patchify(img, patches, gray, gray_split=False):
    _chk(img, torch.float32, "patchify.img"); _chk(patches, torch.bfloat16, "patchify.patches")
    _chk(gray, torch.bfloat16, "patchify.gray")
    assert img.is_contiguous() and img.dim() == 4 and tuple(img.shape[2:]) == (1024, 1024), img.shape
    B, Cc = img.shape[0], img.shape[1]
    assert patches.is_contiguous() and patches.numel() == B * 4096 * 256 * Cc
    assert gray is None or (Cc == 3 and gray.is_contiguous() and gray.numel() == B * 1024 * 1024 * (3 if gray_split else 1))
    _lib.call("wm_patchify", img.data_ptr(), patches.data_ptr(), _ptr(gray), int(bool(gray_split)), B, Cc, _stream())


_define("patchify(Tensor img, Tensor(a!) patches, Tensor(b!)? gray, bool gray_split=False) -> ()", _patchify)


def _transpose(x, out):
    assert x.is_contiguous() and out.is_contiguous() and x.dim() == 3 and x.dtype == out.dtype
    b, R, Cc = x.shape
    assert out.numel() == x.numel()
    _lib.call("wm_transpose", x.data_ptr(), out.data_ptr(), b, R, Cc, x.element_size(), _stream())


_define("transpose(Tensor x, Tensor(a!) out) -> ()", _transpose)


def _transpose_split(x, out):
    _chk(x, torch.float32, "transpose_split.x"); _chk(out, torch.bfloat16, "transpose_split.out")
    assert x.is_contiguous() and out.is_contiguous() and x.dim() == 3
    b, R, Cc = x.shape
    assert out.numel() == 3 * x.numel()
    _lib.call("wm_transpose_split", x.data_ptr(), out.data_ptr(), b, R, Cc, _stream())


_define("transpose_split(Tensor x, Tensor(a!) out) -> ()", _transpose_split)


def _hfc_finalize(img, low_t, patches, hfc_img):
    _chk(img, torch.float32, "hfc_finalize.img"); _chk(low_t, torch.float32, "hfc_finalize.low_t")
    _chk(patches, torch.bfloat16, "hfc_finalize.patches"); _chk(hfc_img, torch.float32, "hfc_finalize.hfc_img")
    B = img.shape[0]
    assert img.is_contiguous() and low_t.is_contiguous() and low_t.numel() == B * 1024 * 1024
    assert patches.is_contiguous() and patches.numel() == B * 4096 * 256
    _lib.call("wm_hfc_finalize", img.data_ptr(), low_t.data_ptr(), patches.data_ptr(), _ptr(hfc_img), B, _stream())


_define("hfc_finalize(Tensor img, Tensor low_t, Tensor(a!) patches, Tensor(b!)? hfc_img) -> ()", _hfc_finalize)


def _add_cast(a, b, b_mod, out):
    _chk(a, torch.float32, "add_cast.a"); _chk(b, torch.float32, "add_cast.b"); _chk(out, torch.bfloat16, "add_cast.out")
    assert out.is_contiguous() and (a is None or (a.is_contiguous() and out.numel() == a.numel())) and (a is not None or b is not None)
    D = out.shape[-1]
    _lib.call("wm_add_cast", _ptr(a), _ptr(b), int(b_mod), out.data_ptr(), out.numel() // D, D, _stream())


_define("add_cast(Tensor? a, Tensor? b, int b_mod, Tensor(a!) out) -> ()", _add_cast)


# ------------------------------------------------------------------ attention
def _attn_flash(q, q_col0, k, k_col0, v, v_col0, rel_table, out, B, H, Tq, Tk, hd, scale):
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t, torch.bfloat16, "attn_flash." + n)
        assert t.dim() == 2
    _chk(rel_table, torch.bfloat16, "attn_flash.rel_table")
    assert rel_table is None or (rel_table.is_contiguous() and tuple(rel_table.shape) == (256, hd))
    assert out.shape[0] >= B * Tq and out.shape[1] >= H * hd
    _lib.call("wm_attn_flash", q.data_ptr(), q.shape[0], q.shape[1], q.stride(0), int(q_col0),
              k.data_ptr(), k.shape[0], k.shape[1], k.stride(0), int(k_col0),
              v.data_ptr(), v.shape[0], v.shape[1], v.stride(0), int(v_col0),
              _ptr(rel_table), out.data_ptr(), out.stride(0), B, H, Tq, Tk, hd, float(scale), _stream())


_define("attn_flash(Tensor q, int q_col0, Tensor k, int k_col0, Tensor v, int v_col0, Tensor? rel_table, "
        "Tensor(a!) out, int B, int H, int Tq, int Tk, int hd, float scale) -> ()", _attn_flash)


def _attn_window(qkv, rel_table, out, H, scale):
    _chk(qkv, torch.bfloat16, "attn_window.qkv"); _chk(rel_table, torch.bfloat16, "attn_window.rel_table")
    _chk(out, torch.bfloat16, "attn_window.out")
    assert qkv.is_contiguous() and out.is_contiguous() and rel_table.is_contiguous()
    D = out.shape[-1]
    B = out.numel() // (4096 * D)
    assert qkv.numel() == B * 4096 * 3 * D and tuple(rel_table.shape) == (64, D // H)
    _lib.call("wm_attn_window", qkv.data_ptr(), rel_table.data_ptr(), out.data_ptr(), B, H, D, float(scale), _stream())


_define("attn_window(Tensor qkv, Tensor rel_table, Tensor(a!) out, int H, float scale) -> ()", _attn_window)


def _attn_small(q, k, v, out, B, H, Tq, Tk, hd, scale):
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t, torch.bfloat16, "attn_small." + n)
        assert t.dim() == 2
    _lib.call("wm_attn_small", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
              out.data_ptr(), out.stride(0), B, H, Tq, Tk, hd, float(scale), _stream())


_define("attn_small(Tensor q, Tensor k, Tensor v, Tensor(a!) out, int B, int H, int Tq, int Tk, int hd, "
        "float scale) -> ()", _attn_small)


# ------------------------------------------------------------------ post-process
def _postprocess(logits, boxes, sizes, thr, from_prob, packed, query_idx, labels, counts):
    _chk(logits, torch.float32, "postprocess.logits"); _chk(boxes, torch.float32, "postprocess.boxes")
    _chk(sizes, torch.int64, "postprocess.sizes"); _chk(packed, torch.float32, "postprocess.packed")
    _chk(query_idx, torch.int32, "postprocess.query_idx"); _chk(counts, torch.int32, "postprocess.counts")
    _chk(labels, torch.int64, "postprocess.labels")
    assert labels is None or (labels.is_contiguous() and labels.numel() == logits.shape[0] * logits.shape[1])
    assert logits.is_contiguous() and boxes.is_contiguous() and sizes.is_contiguous() and packed.is_contiguous()
    B, Q, C1 = logits.shape
    assert tuple(boxes.shape) == (B, Q, 4) and tuple(sizes.shape) == (B, 2) and packed.numel() == B * Q * 6
    _lib.call("wm_postprocess", logits.data_ptr(), boxes.data_ptr(), sizes.data_ptr(), float(thr), int(from_prob),
              packed.data_ptr(), query_idx.data_ptr(), _ptr(labels), counts.data_ptr(), B, Q, C1, _stream())


_define("postprocess(Tensor logits, Tensor boxes, Tensor sizes, float thr, int from_prob, Tensor(a!) packed, "
        "Tensor(b!) query_idx, Tensor(c!)? labels, Tensor(d!) counts) -> ()", _postprocess)


def _sigmoid_topk(logits, boxes, prob_ws, order_ws, scores, labels, query, out_boxes, C, K, from_prob):
    B, Q, C1 = logits.shape
    for t in (logits, boxes, prob_ws, order_ws, scores, labels, query, out_boxes):
        assert t.is_cuda and t.is_contiguous()
    assert prob_ws.numel() == B * Q * C and order_ws.numel() == B * Q * C
    _lib.call("wm_sigmoid_topk", logits.data_ptr(), boxes.data_ptr(), prob_ws.data_ptr(), order_ws.data_ptr(),
              scores.data_ptr(), labels.data_ptr(), query.data_ptr(), out_boxes.data_ptr(), B, Q, C1, int(C), int(K),
              int(from_prob), _stream())


_define("sigmoid_topk(Tensor logits, Tensor boxes, Tensor(a!) prob_ws, Tensor(b!) order_ws, Tensor(c!) scores, "
        "Tensor(d!) labels, Tensor(e!) query, Tensor(f!) out_boxes, int C, int K, int from_prob) -> ()", _sigmoid_topk)


def _nms(boxes, scores, labels, iou_thr, order_ws, mask_ws, keep, num_keep):
    _chk(boxes, torch.float32, "nms.boxes"); _chk(scores, torch.float32, "nms.scores"); _chk(labels, torch.int64, "nms.labels")
    _chk(order_ws, torch.int32, "nms.order_ws"); _chk(mask_ws, torch.int64, "nms.mask_ws")
    _chk(keep, torch.int64, "nms.keep"); _chk(num_keep, torch.int32, "nms.num_keep")
    n = boxes.shape[0]
    assert boxes.is_contiguous() and scores.is_contiguous() and (labels is None or labels.is_contiguous())
    assert mask_ws.numel() >= n * ((n + 63) // 64) and order_ws.numel() >= n and keep.numel() >= n
    _lib.call("wm_nms", boxes.data_ptr(), scores.data_ptr(), _ptr(labels), n, float(iou_thr), order_ws.data_ptr(),
              mask_ws.data_ptr(), keep.data_ptr(), num_keep.data_ptr(), _stream())


_define("nms(Tensor boxes, Tensor scores, Tensor? labels, float iou_thr, Tensor(a!) order_ws, Tensor(b!) mask_ws, "
        "Tensor(c!) keep, Tensor(d!) num_keep) -> ()", _nms)

def _nms_batched(packed, counts, score_thr, iou_thr, per_class, keep_idx, keep_cnt):
    _chk(packed, torch.float32, "nms_batched.packed"); _chk(counts, torch.int32, "nms_batched.counts")
    _chk(keep_idx, torch.int32, "nms_batched.keep_idx"); _chk(keep_cnt, torch.int32, "nms_batched.keep_cnt")
    B, Q, six = packed.shape
    assert six == 6 and packed.is_contiguous() and keep_idx.numel() == B * Q and keep_cnt.numel() == B
    _lib.call("wm_nms_batched", packed.data_ptr(), counts.data_ptr(), B, Q, float(score_thr), float(iou_thr),
              int(per_class), keep_idx.data_ptr(), keep_cnt.data_ptr(), _stream())


_define("nms_batched(Tensor packed, Tensor counts, float score_thr, float iou_thr, int per_class, "
        "Tensor(a!) keep_idx, Tensor(b!) keep_cnt) -> ()", _nms_batched)

# ------------------------------------------------------------------ tile front-end / merge / COCO packing (section 8f)
def _tiles_from_u8(img, origins, content_h, content_w, mean, std, out):
    _chk(img, torch.uint8, "tiles_from_u8.img"); _chk(origins, torch.int32, "tiles_from_u8.origins")
    _chk(out, torch.float32, "tiles_from_u8.out")
    H, W, three = img.shape
    T = origins.shape[0]
    assert three == 3 and img.stride(1) == 3 and origins.is_contiguous() and out.is_contiguous()
    assert tuple(out.shape) == (T, 3, 1024, 1024) and len(mean) == 3 and len(std) == 3
    import ctypes as C
    # fp32 mean / std exactly as torch.as_tensor(mean, dtype=float32) rounds them (torchvision normalize)
    m = (C.c_float * 3)(*[float(x) for x in mean]); sd = (C.c_float * 3)(*[float(x) for x in std])
    _lib.call("wm_tiles_from_u8", img.data_ptr(), H, W, img.stride(0), origins.data_ptr(), T, int(content_h), int(content_w),
              C.addressof(m), C.addressof(sd), out.data_ptr(), _stream())


_define("tiles_from_u8(Tensor img, Tensor origins, int content_h, int content_w, float[] mean, float[] std, "
        "Tensor(a!) out) -> ()", _tiles_from_u8)


def _resize_tiles_u8(img, origins, tile_h, tile_w, tmp, out, xbounds, xk, ybounds, yk):
    _chk(img, torch.uint8, "resize_tiles_u8.img"); _chk(origins, torch.int32, "resize_tiles_u8.origins")
    _chk(tmp, torch.uint8, "resize_tiles_u8.tmp"); _chk(out, torch.uint8, "resize_tiles_u8.out")
    for t, n in ((xbounds, "xbounds"), (xk, "xk"), (ybounds, "ybounds"), (yk, "yk")):
        _chk(t, torch.int32, "resize_tiles_u8." + n)
        assert t.is_contiguous()
    H, W, three = img.shape
    T, oh, ow, _ = out.shape
    assert three == 3 and img.stride(1) == 3 and origins.is_contiguous() and origins.shape[0] == T
    assert out.is_contiguous() and tmp.is_contiguous() and tmp.numel() >= T * tile_h * ow * 3
    assert tuple(xbounds.shape) == (ow, 2) and xk.shape[0] == ow and tuple(ybounds.shape) == (oh, 2) and yk.shape[0] == oh
    _lib.call("wm_resize_tiles_u8", img.data_ptr(), H, W, img.stride(0), origins.data_ptr(), T, int(tile_h), int(tile_w),
              tmp.data_ptr(), out.data_ptr(), oh, ow, xbounds.data_ptr(), xk.data_ptr(), xk.shape[1], ybounds.data_ptr(),
              yk.data_ptr(), yk.shape[1], _stream())


_define("resize_tiles_u8(Tensor img, Tensor origins, int tile_h, int tile_w, Tensor(a!) tmp, Tensor(b!) out, Tensor xbounds, "
        "Tensor xk, Tensor ybounds, Tensor yk) -> ()", _resize_tiles_u8)


def _merge_detections(packed, counts, origins, score_thr, tile_n_ws, boxes, scores, labels, src, total):
    _chk(packed, torch.float32, "merge_detections.packed"); _chk(counts, torch.int32, "merge_detections.counts")
    _chk(origins, torch.int32, "merge_detections.origins"); _chk(tile_n_ws, torch.int32, "merge_detections.tile_n_ws")
    _chk(boxes, torch.float32, "merge_detections.boxes"); _chk(scores, torch.float32, "merge_detections.scores")
    _chk(labels, torch.int64, "merge_detections.labels"); _chk(src, torch.int32, "merge_detections.src")
    _chk(total, torch.int32, "merge_detections.total")
    T, Q, six = packed.shape
    assert six == 6 and packed.is_contiguous() and counts.numel() == T and origins.numel() == 2 * T
    assert boxes.numel() >= 4 * T * Q and scores.numel() >= T * Q and labels.numel() >= T * Q and src.numel() >= 2 * T * Q
    assert tile_n_ws.numel() >= T
    _lib.call("wm_merge_detections", packed.data_ptr(), counts.data_ptr(), origins.data_ptr(), T, Q, float(score_thr),
              tile_n_ws.data_ptr(), boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), src.data_ptr(),
              total.data_ptr(), _stream())


_define("merge_detections(Tensor packed, Tensor counts, Tensor origins, float score_thr, Tensor(a!) tile_n_ws, "
        "Tensor(b!) boxes, Tensor(c!) scores, Tensor(d!) labels, Tensor(e!) src, Tensor(f!) total) -> ()",
        _merge_detections)


def _pack_coco(boxes, scores, labels, keep, n_keep, out_xywh_score, out_category):
    _chk(boxes, torch.float32, "pack_coco.boxes"); _chk(scores, torch.float32, "pack_coco.scores")
    _chk(labels, torch.int64, "pack_coco.labels"); _chk(keep, torch.int64, "pack_coco.keep")
    _chk(out_xywh_score, torch.float32, "pack_coco.out"); _chk(out_category, torch.int64, "pack_coco.out_category")
    assert boxes.is_contiguous() and out_xywh_score.is_contiguous()
    assert out_xywh_score.numel() >= 5 * n_keep and out_category.numel() >= n_keep
    assert keep is None or keep.numel() >= n_keep
    _lib.call("wm_pack_coco", boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), _ptr(keep), int(n_keep),
              out_xywh_score.data_ptr(), out_category.data_ptr(), _stream())


_define("pack_coco(Tensor boxes, Tensor scores, Tensor labels, Tensor? keep, int n_keep, Tensor(a!) out_xywh_score, "
        "Tensor(b!) out_category) -> ()", _pack_coco)

# ------------------------------------------------------------------ evaluation-time criterion (section 8f row 4)
def _match_cost(logits, boxes, tgt_ids, tgt_boxes, w_class, w_bbox, w_giou, cost):
    _chk(logits, torch.float32, "match_cost.logits"); _chk(boxes, torch.float32, "match_cost.boxes")
    _chk(tgt_ids, torch.int64, "match_cost.tgt_ids"); _chk(tgt_boxes, torch.float32, "match_cost.tgt_boxes")
    _chk(cost, torch.float32, "match_cost.cost")
    rows, C1 = logits.shape
    T = tgt_ids.shape[0]
    assert logits.is_contiguous() and boxes.is_contiguous() and tgt_boxes.is_contiguous() and cost.is_contiguous()
    assert tuple(boxes.shape) == (rows, 4) and tuple(tgt_boxes.shape) == (T, 4) and tuple(cost.shape) == (rows, T)
    _lib.call("wm_match_cost", logits.data_ptr(), boxes.data_ptr(), tgt_ids.data_ptr(), tgt_boxes.data_ptr(), rows, T, C1,
              float(w_class), float(w_bbox), float(w_giou), cost.data_ptr(), _stream())


_define("match_cost(Tensor logits, Tensor boxes, Tensor tgt_ids, Tensor tgt_boxes, float w_class, float w_bbox, float w_giou, "
        "Tensor(a!) cost) -> ()", _match_cost)


def _set_criterion(logits, boxes, m_row, m_label, m_box, tgt_len, empty_weight, num_boxes, tcls_ws, out5):
    _chk(logits, torch.float32, "set_criterion.logits"); _chk(boxes, torch.float32, "set_criterion.boxes")
    _chk(m_row, torch.int32, "set_criterion.m_row"); _chk(m_label, torch.int64, "set_criterion.m_label")
    _chk(m_box, torch.float32, "set_criterion.m_box"); _chk(tgt_len, torch.int32, "set_criterion.tgt_len")
    _chk(empty_weight, torch.float32, "set_criterion.empty_weight"); _chk(tcls_ws, torch.int32, "set_criterion.tcls_ws")
    _chk(out5, torch.float32, "set_criterion.out5")
    B, Q, C1 = logits.shape
    n = m_row.shape[0]
    assert logits.is_contiguous() and boxes.is_contiguous() and tuple(boxes.shape) == (B, Q, 4)
    assert m_label.shape[0] == n and m_box.numel() == 4 * n and m_box.is_contiguous() and tgt_len.numel() == B
    assert empty_weight.numel() == C1 and tcls_ws.numel() >= B * Q and out5.numel() == 5
    _lib.call("wm_set_criterion", logits.data_ptr(), boxes.data_ptr(), m_row.data_ptr(), m_label.data_ptr(), m_box.data_ptr(), n,
              tgt_len.data_ptr(), empty_weight.data_ptr(), B, Q, C1, float(num_boxes), tcls_ws.data_ptr(), out5.data_ptr(),
              _stream())


_define("set_criterion(Tensor logits, Tensor boxes, Tensor m_row, Tensor m_label, Tensor m_box, Tensor tgt_len, "
        "Tensor empty_weight, float num_boxes, Tensor(a!) tcls_ws, Tensor(b!) out5) -> ()", _set_criterion)

ops = torch.ops.wm_b200
