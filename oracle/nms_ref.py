"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the yolo-re
post-processing: ``non_max_suppression`` (src/yolo/utils/nms.py:19-94) on top of greedy NMS.

The greedy step itself is third-party in the reference: ``torchvision.ops.nms``
(optional extra, pinned 0.24.1 in uv.lock:2430; this image has 0.26.0), called at
src/yolo/utils/nms.py:99-102, with the in-tree ``_nms_pure``/``_box_iou``
(nms.py:107-152) as its semantic restatement.  torchvision's published algorithm, restated
here: stable descending sort of the scores; walk the sorted list; a box is kept iff no
previously *kept* box has ``double(inter / (area_i + area_j - inter)) > iou_thres`` (fp32 IoU
compared against the threshold as a double -- verified against torchvision's CPU kernel) with
``inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1))`` -- all in fp32.

Two implementations with identical results:
  * ``nms_numpy``  -- numpy fp32, used for small cases and to check the C port;
  * ``nms_c``      -- oracle/nms_ref.c through ctypes (built by oracle/Makefile), used for
                      the big cases (33 600 candidates per image).

Parity pinning: the reference has NO test of its NMS, so the pin is
the reference itself run in the build container (tests/golden/make_golden.py writes its
detections into tests/golden/) plus ``torchvision.ops.nms`` on the machine running the tests.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def greedy_nms_numpy(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """torchvision.ops.nms semantics (see module docstring).  boxes [n,4] xyxy fp32."""
    boxes = np.ascontiguousarray(boxes, np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    thr = float(iou_thres)            # torchvision's CPU kernel compares the fp32 IoU against the DOUBLE threshold
    dead = np.zeros(len(boxes), bool)
    keep = []
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / (area[i] + area[rest] - inter)
        dead[rest[iou.astype(np.float64) > thr]] = True
    return np.asarray(keep, np.int64)


def nms_image_numpy(pred: np.ndarray, conf_thres=0.25, iou_thres=0.45, max_det=300, classes=None,
                    agnostic=False):
    """One image of non_max_suppression (nms.py:46-92).  pred [A, 4+nc] fp32.
    Returns (det [n,6] fp32, keep_anchor [n] int64)."""
    pred = np.ascontiguousarray(pred, np.float32)
    cls_scores = pred[:, 4:]
    cls_idx = cls_scores.argmax(1)                       # first max on ties (torch.max(dim=1), nms.py:54)
    conf = cls_scores[np.arange(len(pred)), cls_idx]
    mask = conf > np.float32(conf_thres)                 # strict > (nms.py:57)
    if classes is not None:
        mask &= np.isin(cls_idx, np.asarray(classes))    # nms.py:58-59
    cand = np.nonzero(mask)[0]
    if len(cand) == 0:                                   # nms.py:68-70
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    xywh = pred[cand, :4]
    half_w, half_h = xywh[:, 2] / np.float32(2), xywh[:, 3] / np.float32(2)
    boxes = np.stack((xywh[:, 0] - half_w, xywh[:, 1] - half_h,   # xywh2xyxy, nms.py:9-16
                      xywh[:, 0] + half_w, xywh[:, 1] + half_h), 1).astype(np.float32)
    c, k = conf[cand], cls_idx[cand]
    if agnostic:
        nms_boxes = boxes
    else:                                                # class-offset trick, nms.py:79-81 (two rounded fp32 ops)
        off = k.astype(np.float32) * (boxes.max() + np.float32(1))
        nms_boxes = boxes + off[:, None]
    keep = greedy_nms_numpy(nms_boxes, c, iou_thres)[:max_det]
    det = np.concatenate((boxes[keep], c[keep, None], k[keep, None].astype(np.float32)), 1)
    return det.astype(np.float32), cand[keep].astype(np.int64)


def _lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "libnms_ref.so"
        if not so.exists():
            raise FileNotFoundError(f"{so} missing -- run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(str(so))
        _LIB.yre_oracle_nms_image.restype = ctypes.c_int
        _LIB.yre_oracle_nms_image.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def nms_image_c(pred: np.ndarray, conf_thres=0.25, iou_thres=0.45, max_det=300, classes=None, agnostic=False):
    pred = np.ascontiguousarray(pred, np.float32)
    a, ch = pred.shape
    det = np.zeros((max_det, 6), np.float32)
    keep = np.zeros((max_det,), np.int64)
    cl = np.ascontiguousarray(classes if classes is not None else [], np.int32)
    n = _lib().yre_oracle_nms_image(pred.ctypes.data, a, ch - 4, conf_thres, iou_thres, max_det,
                                    cl.ctypes.data if classes is not None else None,
                                    len(cl) if classes is not None else -1, int(bool(agnostic)),
                                    det.ctypes.data, keep.ctypes.data)
    if n < 0:
        raise MemoryError("oracle nms: allocation failed")
    return det[:n].copy(), keep[:n].copy()


def non_max_suppression(predictions, conf_thres=0.25, iou_thres=0.45, max_det=300, classes=None,
                        agnostic=False, impl: str = "auto", return_keep: bool = False):
    """Batch wrapper with the reference signature (nms.py:19-26).  predictions [B,A,4+nc]
    (numpy or torch CPU).  Returns list of [n,6] float32 numpy arrays (and keep-anchor lists)."""
    if hasattr(predictions, "detach"):
        predictions = predictions.detach().cpu().float().numpy()
    if impl == "auto":
        impl = "c" if (_HERE / "_build" / "libnms_ref.so").exists() else "numpy"
    fn = nms_image_c if impl == "c" else nms_image_numpy
    dets, keeps = [], []
    for i in range(predictions.shape[0]):
        d, k = fn(predictions[i], conf_thres, iou_thres, max_det, classes, agnostic)
        dets.append(d)
        keeps.append(k)
    return (dets, keeps) if return_keep else dets
