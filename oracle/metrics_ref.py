"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's detection metrics.

  box_iou      src/yolo/eval/metrics.py:10-31   (fp32, same operation order)
  compute_ap   src/yolo/eval/metrics.py:34-60   (101-point interpolation, float64)
  compute_map  src/yolo/eval/metrics.py:63-198  (per class: global stable score sort, greedy match of each prediction to its
                                                 best-IoU ground truth of the same image and class, at every IoU threshold)

Numpy only.  Semantics that matter for bit-equality: IoU is fp32 `inter / ((area1 + area2) - inter)`; the best ground truth
is the FIRST arg-max; `best_iou >= thr` is evaluated in fp32 (torch casts the Python threshold to the tensor's dtype);
Python's `sort(reverse=True)` is stable, i.e. ties keep (image, detection) order.
Pinned against the reference's compute_map on seeded random cases (tests/golden/metrics_cases.npz)."""
from __future__ import annotations

import numpy as np


def box_iou(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    b1, b2 = b1.astype(np.float32), b2.astype(np.float32)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, np.float32(0), None)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = (area1[:, None] + area2[None, :]) - inter
    return inter / union


def compute_ap(recall: np.ndarray, precision: np.ndarray) -> float:
    """COCO-style AP: precision envelope (running maximum from the right, with the (0,1) and (1,0) sentinels) sampled at
    the 101 recall levels 0, 0.01, ..., 1 -- at each level the envelope value of the first point whose recall reaches it."""
    import bisect
    rec = [0.0] + [float(v) for v in recall] + [1.0]
    env = [1.0] + [float(v) for v in precision] + [0.0]
    for k in reversed(range(len(env) - 1)):
        if env[k + 1] > env[k]:
            env[k] = env[k + 1]
    samples = []
    for level in np.linspace(0, 1, 101):
        k = bisect.bisect_left(rec, level)             # recall is non-decreasing
        samples.append(env[k] if k < len(rec) else 0.0)
    return float(np.asarray(samples, np.float64).mean())


def match_image(pred_boxes, pred_classes, gt_boxes, gt_classes, thresholds):
    """TP flags [n, T] of one image whose predictions are already in matching order (descending score, stable)."""
    n, T = len(pred_boxes), len(thresholds)
    tp = np.zeros((n, T), np.uint8)
    thr32 = np.asarray(thresholds, np.float64).astype(np.float32)
    matched = np.zeros((T, len(gt_boxes)), bool)
    for d in range(n):
        sel = np.nonzero(gt_classes == pred_classes[d])[0]
        if len(sel) == 0:
            continue
        ious = box_iou(pred_boxes[d:d + 1], gt_boxes[sel])[0]
        j = int(np.argmax(ious))                    # first maximum
        for t in range(T):
            if ious[j] >= thr32[t] and not matched[t, sel[j]]:
                tp[d, t] = 1
                matched[t, sel[j]] = True
    return tp


def compute_map(pred_boxes, pred_scores, pred_classes, gt_boxes, gt_classes, num_classes, iou_thresholds=None):
    if iou_thresholds is None:
        iou_thresholds = [0.5 + 0.05 * i for i in range(10)]
    n_img = len(pred_boxes)
    # per image: order predictions by descending score (stable) -- within an image this is the order the global sort visits them
    orders = [np.argsort(-np.asarray(s, np.float32), kind="stable") for s in pred_scores]
    tps = [match_image(np.asarray(pred_boxes[i], np.float32)[orders[i]], np.asarray(pred_classes[i])[orders[i]],
                       np.asarray(gt_boxes[i], np.float32).reshape(-1, 4), np.asarray(gt_classes[i]), iou_thresholds) for i in range(n_img)]
    all_aps = {t: [] for t in iou_thresholds}
    for c in range(num_classes):
        total_gt = int(sum(int((np.asarray(g) == c).sum()) for g in gt_classes))
        if total_gt == 0:
            continue
        sc, flags = [], []
        for i in range(n_img):
            m = np.asarray(pred_classes[i])[orders[i]] == c
            if m.any():
                sc.append(np.asarray(pred_scores[i], np.float32)[orders[i]][m])
                flags.append(tps[i][m])
        if not sc:
            for t in iou_thresholds:
                all_aps[t].append(0.0)
            continue
        sc, flags = np.concatenate(sc), np.concatenate(flags)
        order = np.argsort(-sc, kind="stable")
        flags = flags[order]
        for ti, t in enumerate(iou_thresholds):
            tp = flags[:, ti].astype(np.float64)
            tpc, fpc = np.cumsum(tp), np.cumsum(1.0 - tp)
            all_aps[t].append(compute_ap(tpc / total_gt, tpc / (tpc + fpc)))
    res = {"map50": float(np.mean(all_aps[0.5])) if 0.5 in all_aps and all_aps[0.5] else 0.0,
           "map75": float(np.mean(all_aps[0.75])) if 0.75 in all_aps and all_aps[0.75] else 0.0}
    vals = []
    for t in iou_thresholds:
        vals.extend(all_aps.get(t, []))
    res["map"] = float(np.mean(vals)) if vals else 0.0
    return res
