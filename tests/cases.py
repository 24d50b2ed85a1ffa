"""Seeded synthetic NMS inputs shared by the golden generator and the tests
(distribution from SURVEY.md section 8d: one dominant class per anchor, others <= 9e-4)."""
import torch


def synth_pred(B, A, nc, S, seed, quant=None, neg=False, wide=False):
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros(B, A, 4 + nc)
    p[..., 0:2] = torch.rand(B, A, 2, generator=g) * S
    p[..., 2:4] = torch.rand(B, A, 2, generator=g) * ((0.6 if wide else 0.2) * S) + 4
    if neg:                       # boxes hanging over the top/left edge -> negative x1/y1
        p[..., 0:2] -= 0.15 * S
    p[..., 4:] = torch.rand(B, A, nc, generator=g) * 9e-4
    dom = torch.randint(0, nc, (B, A), generator=g)
    sc = torch.rand(B, A, generator=g)
    if quant:                     # heavy score ties
        sc = torch.round(sc * quant) / quant
    p.scatter_(2, (dom + 4).unsqueeze(-1), sc.unsqueeze(-1))
    return p


NMS_CASES = {
    "plain":      dict(gen=dict(B=2, A=2000, nc=80, S=640, seed=3), kw=dict(conf_thres=0.25, iou_thres=0.45)),
    "ties":       dict(gen=dict(B=2, A=3000, nc=80, S=640, seed=4, quant=64), kw=dict(conf_thres=0.25, iou_thres=0.45)),
    "few_cls_neg": dict(gen=dict(B=2, A=3000, nc=3, S=640, seed=5, quant=16, neg=True), kw=dict(conf_thres=0.05, iou_thres=0.45)),
    "stress":     dict(gen=dict(B=1, A=8400, nc=80, S=640, seed=6, neg=True), kw=dict(conf_thres=0.001, iou_thres=0.6)),
    "agnostic":   dict(gen=dict(B=2, A=1500, nc=80, S=640, seed=7, quant=32), kw=dict(conf_thres=0.25, iou_thres=0.45, agnostic=True)),
    "classes":    dict(gen=dict(B=2, A=1500, nc=80, S=640, seed=8), kw=dict(conf_thres=0.25, iou_thres=0.45, classes=[1, 5, 7])),
    "empty":      dict(gen=dict(B=2, A=500, nc=80, S=640, seed=9), kw=dict(conf_thres=0.99999, iou_thres=0.45)),
    "iou0":       dict(gen=dict(B=1, A=1000, nc=4, S=64, seed=10, quant=8, neg=True), kw=dict(conf_thres=0.1, iou_thres=0.0)),
    "iou1":       dict(gen=dict(B=1, A=1000, nc=4, S=64, seed=11, quant=8, neg=True), kw=dict(conf_thres=0.1, iou_thres=1.0)),
    "small_maxdet": dict(gen=dict(B=2, A=1200, nc=20, S=320, seed=12, wide=True), kw=dict(conf_thres=0.2, iou_thres=0.5, max_det=17)),
    "few_keep":   dict(gen=dict(B=3, A=900, nc=2, S=96, seed=13, wide=True), kw=dict(conf_thres=0.3, iou_thres=0.3)),
    "at_thr_0.6": dict(gen="at_threshold", kw=dict(conf_thres=0.25, iou_thres=0.6)),
    "at_thr_0.45": dict(gen="at_threshold", kw=dict(conf_thres=0.25, iou_thres=0.45)),
    "at_thr_agn": dict(gen="at_threshold", kw=dict(conf_thres=0.25, iou_thres=0.6, agnostic=True)),
}


def make_pred(case):
    return at_threshold_pred() if case["gen"] == "at_threshold" else synth_pred(**case["gen"])


def at_threshold_pred(nc=4):
    """Pairs whose fp32 IoU equals float32(thr) exactly, with float32(thr) > thr (thr=0.6) or
    < thr (thr=0.45): the reference (torchvision CPU) compares the fp32 IoU with the threshold as a
    DOUBLE, so the 0.6 pair is suppressed although `iou > float32(0.6)` is false."""
    import numpy as np
    rows = []
    for k, thr in enumerate((0.6, 0.45, 0.6, 0.45)):
        t = float(np.float32(thr))
        x0 = 100.0 * k
        # box A: unit square scaled by 32; box B: same height, width t -> inter = t*A, union = A
        for w, sc in ((32.0, 0.9 - 0.01 * k), (32.0 * t, 0.8 - 0.01 * k)):
            row = [x0 + w / 2, 16.0, w, 32.0] + [0.0] * nc
            row[4 + (k % nc)] = sc
            rows.append(row)
    return torch.tensor([rows], dtype=torch.float32)


# ---- pre/post-processing cases: name -> (h, w, new_shape, seed) ---------------------------
PREPROC_CASES = {
    "vga": (480, 640, 640, 1),            # width already 640: no resize, pad top/bottom
    "up_500x375": (375, 500, 640, 2),     # up-scaling
    "hd": (1080, 1920, 640, 3),           # down-scaling, non-integer ratio
    "area2x": (1280, 960, 640, 4),        # exact 2x down-scale -> cv2 switches to INTER_AREA
    "odd": (333, 517, 640, 5),
    "tiny_up": (100, 80, 640, 6),         # 6.4x up-scaling, many clipped border rows
    "square": (640, 640, 640, 7),         # identity
    "portrait_1280": (1280, 720, 1280, 8),
    "odd_pad": (481, 641, 640, 9),        # odd padding -> top != bottom
}


def preproc_image(h: int, w: int, seed: int):
    """Smooth-ish synthetic BGR image with full 0..255 range (seeded, numpy only)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = np.repeat(np.repeat(base, 8, 0), 8, 1)[:h, :w].astype(np.int32)
    img += rng.integers(-24, 25, (h, w, 3), dtype=np.int32)
    return np.clip(img, 0, 255).astype(np.uint8)


def preproc_boxes(S: int, seed: int, n: int = 64):
    """xyxy boxes in letterboxed-input pixels, some outside the un-padded image so the clip matters."""
    import numpy as np
    rng = np.random.default_rng(1000 + seed)
    c = rng.uniform(-20, S + 20, (n, 2)); wh = rng.uniform(2, S / 2, (n, 2))
    return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)


# ---- detection-metric cases: name -> (images, classes, max detections/img, max gts/img, seed) --------
METRIC_CASES = {
    "small": (4, 3, 12, 6, 1),
    "coco_like": (16, 80, 300, 40, 2),
    "ties": (6, 2, 40, 10, 3),            # quantised scores (ties across images) and boxes snapped to a grid (IoU ties)
    "empty_mix": (8, 5, 20, 8, 4),        # some images without detections, some without ground truth
}


def metric_case(name: str):
    """(pred_boxes, pred_scores, pred_classes, gt_boxes, gt_classes, num_classes) as lists of numpy arrays; detections
    per image in descending-score order (what NMS returns)."""
    import numpy as np
    n_img, nc, max_det, max_gt, seed = METRIC_CASES[name]
    rng = np.random.default_rng(seed)
    pb, ps, pc, gb, gc = [], [], [], [], []
    for i in range(n_img):
        m = int(rng.integers(0, max_gt + 1)) if not (name == "empty_mix" and i % 3 == 0) else 0
        c = rng.uniform(40, 600, (m, 2)); wh = rng.uniform(10, 200, (m, 2))
        g = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        gcl = rng.integers(0, nc, m).astype(np.int64)
        n = int(rng.integers(0, max_det + 1)) if not (name == "empty_mix" and i % 4 == 1) else 0
        # detections: jittered copies of ground truths (some with the wrong class) plus random boxes
        src = rng.integers(0, max(m, 1), n)
        jit = rng.normal(0, 12, (n, 4)).astype(np.float32)
        d = (g[src] + jit) if m else np.zeros((n, 4), np.float32)
        rnd = rng.random(n) < 0.35
        c2 = rng.uniform(40, 600, (n, 2)); wh2 = rng.uniform(10, 200, (n, 2))
        d = np.where(rnd[:, None] | (m == 0), np.concatenate([c2 - wh2 / 2, c2 + wh2 / 2], 1), d).astype(np.float32)
        d[:, 2:] = np.maximum(d[:, 2:], d[:, :2] + 1)
        dcl = np.where(rng.random(n) < 0.8, gcl[src] if m else rng.integers(0, nc, n), rng.integers(0, nc, n)).astype(np.int64)
        sc = rng.random(n).astype(np.float32)
        if name == "ties":
            sc = (np.round(sc * 8) / 8).astype(np.float32)
            d = (np.round(d / 16) * 16).astype(np.float32); d[:, 2:] = np.maximum(d[:, 2:], d[:, :2] + 16)
            g = (np.round(g / 16) * 16).astype(np.float32); g[:, 2:] = np.maximum(g[:, 2:], g[:, :2] + 16)
        o = np.argsort(-sc, kind="stable")
        pb.append(d[o]); ps.append(sc[o]); pc.append(dcl[o]); gb.append(g); gc.append(gcl)
    return pb, ps, pc, gb, gc, nc
