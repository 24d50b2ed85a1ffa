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
