"""Batched class-aware NMS on the GPU -- drop-in for ``yolo.utils.nms.non_max_suppression``
(src/yolo/utils/nms.py:19-94): same signature, same return type (list of ``[n,6]`` tensors
``[x1,y1,x2,y2,conf,cls]`` per image), bit-identical results given the same predictions."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_ws_cache: dict = {}


def nms_raw(predictions: torch.Tensor, conf_thres=0.25, iou_thres=0.45, max_det=300, classes=None, agnostic=False):
    """Runs K7 and returns the padded device results ``(out[B,max_det,6], counts[B] int32,
    keep_anchor[B,max_det] int64)`` without any host synchronisation."""
    if not predictions.is_cuda:
        raise L.YreError("the yolo-re B200 path runs on CUDA tensors only (there is no CPU fallback)")
    if predictions.dim() != 3 or predictions.shape[2] <= 4:
        raise ValueError("predictions must be [batch, num_anchors, 4 + num_classes]")
    lib = L.lib()
    pred = predictions.contiguous().float()
    Bn, A, ch = pred.shape
    dev = pred.device
    out = torch.empty((Bn, max_det, 6), dtype=torch.float32, device=dev)
    counts = torch.empty((Bn,), dtype=torch.int32, device=dev)
    keep = torch.empty((Bn, max_det), dtype=torch.int64, device=dev)
    if Bn == 0 or A == 0:
        return out, counts.zero_(), keep
    key = (dev.index, Bn, A)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty((lib.yre_nms_workspace_bytes(Bn, A),), dtype=torch.uint8, device=dev)
        _ws_cache.clear()
        _ws_cache[key] = ws
    cls_t = None
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=dev)
    d = L.NmsDesc(pred.data_ptr(), Bn, A, ch - 4, float(conf_thres), float(iou_thres), int(max_det),
                  cls_t.data_ptr() if cls_t is not None and cls_t.numel() else None,
                  (cls_t.numel() if cls_t is not None else -1), int(bool(agnostic)),
                  out.data_ptr(), counts.data_ptr(), keep.data_ptr(), ws.data_ptr(), ws.numel())
    with torch.cuda.device(dev):
        L.check(lib.yre_nms_batched(C.byref(d), torch.cuda.current_stream(dev).cuda_stream), "nms_batched")
    return out, counts, keep


def non_max_suppression(predictions: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        max_det: int = 300, classes: list[int] | None = None, agnostic: bool = False) -> list[torch.Tensor]:
    """predictions: (batch, num_anchors, 4 + num_classes), boxes xywh in pixels, scores already
    sigmoided.  Returns one (n, 6) tensor per image.  The only host sync is reading the B counts."""
    out, counts, _ = nms_raw(predictions, conf_thres, iou_thres, max_det, classes, agnostic)
    n = counts.tolist()
    return [out[i, : n[i]] for i in range(out.shape[0])]
