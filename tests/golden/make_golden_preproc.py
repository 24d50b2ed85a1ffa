"""Regenerates tests/golden/preproc_cases.npz from the REAL reference functions `letterbox` and `scale_boxes`
(scripts/detect.py:40-109) and the reference's pre-processing lines (scripts/detect.py:223-227), run in the build
container where /root/reference and cv2 exist.  The GPU box only reads the committed .npz.

    python tests/golden/make_golden_preproc.py

Inputs are seeded synthetic uint8 BGR images (tests.cases.preproc_image); outputs are stored as a SHA-256 of the
full letterboxed uint8 image plus three probe rows, so the fixture stays small."""
import hashlib
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.modules.setdefault("albumentations", types.ModuleType("albumentations"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))
spec = importlib.util.spec_from_file_location("ref_detect", "/root/reference/scripts/detect.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

from tests.cases import PREPROC_CASES, preproc_image, preproc_boxes  # noqa: E402

out = {}
for name, (h, w, S, seed) in PREPROC_CASES.items():
    img = preproc_image(h, w, seed)
    lb, ratio, pad = ref.letterbox(img, S)
    assert lb.shape == (S, S, 3), (name, lb.shape)
    chw = np.ascontiguousarray(lb[:, :, ::-1].transpose(2, 0, 1))
    x = (torch.from_numpy(chw).float() / 255.0).numpy()              # scripts/detect.py:224-226
    out[f"{name}/sha"] = np.frombuffer(hashlib.sha256(lb.tobytes()).digest(), np.uint8)
    out[f"{name}/rows"] = lb[[0, S // 3, S - 1]]
    out[f"{name}/x_sha"] = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), np.uint8)
    out[f"{name}/ratio"] = np.array(ratio, np.float64)
    out[f"{name}/pad"] = np.array(pad, np.int64)
    boxes = preproc_boxes(S, seed)
    out[f"{name}/boxes_rp"] = ref.scale_boxes(torch.from_numpy(boxes.copy()), (S, S), (h, w), (ratio, pad)).numpy()
    out[f"{name}/boxes_none"] = ref.scale_boxes(torch.from_numpy(boxes.copy()), (S, S), (h, w)).numpy()
np.savez_compressed(Path(__file__).resolve().parent / "preproc_cases.npz", **out)
print("wrote", len(PREPROC_CASES), "cases")
