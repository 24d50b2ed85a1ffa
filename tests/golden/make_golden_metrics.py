"""Regenerates tests/golden/metrics_cases.npz with the REFERENCE's compute_map (src/yolo/eval/metrics.py:63-198) on the
seeded cases of tests.cases.metric_case.  Run in the build container.

    python tests/golden/make_golden_metrics.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.modules.setdefault("albumentations", types.ModuleType("albumentations"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))
from yolo.eval.metrics import compute_map  # noqa: E402  (the reference)
from tests.cases import METRIC_CASES, metric_case  # noqa: E402

out = {}
for name in METRIC_CASES:
    pb, ps, pc, gb, gc, nc = metric_case(name)
    t = lambda xs, dt=None: [torch.from_numpy(x) if dt is None else torch.from_numpy(x).to(dt) for x in xs]
    r = compute_map(t(pb), t(ps), t(pc), t(gb), t(gc), nc)
    r7 = compute_map(t(pb), t(ps), t(pc), t(gb), t(gc), nc, iou_thresholds=[0.7])        # fp32(0.7) < 0.7: the cast matters
    out[name] = np.array([r["map50"], r["map75"], r["map"], r7["map"]], np.float64)
    print(name, out[name])
np.savez(Path(__file__).resolve().parent / "metrics_cases.npz", **out)
