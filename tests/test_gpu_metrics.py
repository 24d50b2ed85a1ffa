"""GPU parity of K11 (yre_match_detections) and the evaluator fast path -- SURVEY.md 8f row 3.  Flag / index work:
bit-exact against the oracle; the final mAP numbers are bit-equal (float64) to what the REFERENCE's compute_map produced."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import metrics_ref as MR
from tests.cases import METRIC_CASES, metric_case

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import DetectionAccumulator, compute_map, match_detections

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "metrics_cases.npz")
THR = [0.5 + 0.05 * i for i in range(10)]


def _t(xs, dev="cuda"):
    return [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in xs]


@pytest.mark.parametrize("name", list(METRIC_CASES))
def test_match_flags_and_map_bit_exact(name):
    pb, ps, pc, gb, gc, nc = metric_case(name)
    dets = [torch.cat([b, s[:, None], c.float()[:, None]], 1) for b, s, c in zip(_t(pb), _t(ps), _t(pc))]
    tp, off = match_detections(dets, _t(gb), _t(gc), THR)
    ref = np.concatenate([MR.match_image(pb[i], pc[i], gb[i], gc[i], THR) for i in range(len(pb))]) if tp.shape[0] else np.zeros((0, 10), np.uint8)
    assert np.array_equal(tp.cpu().numpy(), ref)
    assert off.tolist() == [0] + list(np.cumsum([len(x) for x in pb]))
    r = compute_map(_t(pb), _t(ps), _t(pc), _t(gb), _t(gc), nc)
    r7 = compute_map(_t(pb, "cpu"), _t(ps, "cpu"), _t(pc, "cpu"), _t(gb, "cpu"), _t(gc, "cpu"), nc, iou_thresholds=[0.7])   # CPU lists are moved
    assert np.array_equal(np.array([r["map50"], r["map75"], r["map"], r7["map"]]), GOLD[name])


def test_unsorted_predictions_and_accumulator():
    pb, ps, pc, gb, gc, nc = metric_case("coco_like")
    rng = np.random.default_rng(0)
    perm = [rng.permutation(len(s)) for s in ps]                     # compute_map sorts (stably) itself, like the reference
    r = compute_map(_t([b[p] for b, p in zip(pb, perm)]), _t([s[p] for s, p in zip(ps, perm)]), _t([c[p] for c, p in zip(pc, perm)]),
                    _t(gb), _t(gc), nc)
    ref = MR.compute_map([b[p] for b, p in zip(pb, perm)], [s[p] for s, p in zip(ps, perm)], [c[p] for c, p in zip(pc, perm)], gb, gc, nc)
    assert r == ref
    acc = DetectionAccumulator(nc)
    dets = [torch.cat([b, s[:, None], c.float()[:, None]], 1) for b, s, c in zip(_t(pb), _t(ps), _t(pc))]
    for lo in range(0, len(dets), 5):                                # batches of 5 images, everything stays on the device
        acc.update(dets[lo:lo + 5], _t(gb[lo:lo + 5]), _t(gc[lo:lo + 5]))
    got = acc.compute()
    assert np.array_equal(np.array([got["map50"], got["map75"], got["map"]]), GOLD["coco_like"][:3])
    assert DetectionAccumulator(3).compute() == {"map50": 0.0, "map75": 0.0, "map": 0.0}
    with pytest.raises(yolo_b200.YreError):
        match_detections([torch.zeros((2, 6))], [torch.zeros((1, 4))], [torch.zeros(1)], THR)      # CPU detections: no fallback


def test_evaluator_flow_matches_reference_arithmetic():
    """Evaluator.evaluate() (reference src/yolo/eval/evaluator.py:96-199 flow on the device) == the oracle's compute_map fed
    with the same model's detections and the reference's CPU ground-truth conversion."""
    from oracle import gelan_ref as G
    from yolo_b200 import YOLO, Evaluator, non_max_suppression
    root = Path(__file__).resolve().parents[1]
    nodes, nc = G.load_graph(root / "configs/models/gelan-c.yaml")
    m = YOLO.from_yaml(root / "configs/models/gelan-c.yaml")
    m.load_state_dict(G.calibrated_state_dict(nodes, nc))
    rng = np.random.default_rng(5)
    S, batches = 256, []
    for bi in range(2):
        imgs = G.fractal(3, S, torch.Generator().manual_seed(30 + bi))
        rows = []
        for i in range(3):
            k = int(rng.integers(0, 6))
            cxy, wh = rng.uniform(0.2, 0.8, (k, 2)), rng.uniform(0.05, 0.4, (k, 2))
            rows += [[i, int(rng.integers(0, nc)), *cxy[j], *wh[j]] for j in range(k)]
        batches.append((imgs, torch.tensor(rows, dtype=torch.float32).reshape(-1, 6), None, [(S, S)] * 3))
    got = Evaluator(m, batches, num_classes=nc, conf_thres=0.25, iou_thres=0.45).evaluate()
    # the same flow with the reference's host-side arithmetic and the oracle's metric code
    pb, ps, pc, gb, gc = [], [], [], [], []
    m = m.cuda().eval()
    for imgs, targets, _, _ in batches:
        y, _ = m(imgs.cuda())
        dets = non_max_suppression(y.permute(0, 2, 1).contiguous(), 0.25, 0.45)
        for i in range(imgs.shape[0]):
            d = dets[i].cpu().numpy()
            pb.append(d[:, :4]); ps.append(d[:, 4]); pc.append(d[:, 5].astype(np.int64))
            t = targets[targets[:, 0] == i]
            xywh = t[:, 2:6].clone()
            xywh[:, [0, 2]] *= S; xywh[:, [1, 3]] *= S
            xyxy = torch.zeros_like(xywh)
            xyxy[:, 0] = xywh[:, 0] - xywh[:, 2] / 2; xyxy[:, 1] = xywh[:, 1] - xywh[:, 3] / 2
            xyxy[:, 2] = xywh[:, 0] + xywh[:, 2] / 2; xyxy[:, 3] = xywh[:, 1] + xywh[:, 3] / 2
            gb.append(xyxy.numpy()); gc.append(t[:, 1].long().numpy())
    assert got == MR.compute_map(pb, ps, pc, gb, gc, nc)
    with pytest.raises(NotImplementedError):
        Evaluator(m, batches, debug_dir="/tmp/x")
