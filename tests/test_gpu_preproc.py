"""GPU parity of K8 (letterbox / preprocess) and K9 (scale_boxes) -- SURVEY.md 8f row 1.  Byte and index work:
bit-exact against the oracle restatement (itself pinned to cv2 and to the reference's functions) and against the
committed outputs of the reference (tests/golden/preproc_cases.npz).  fp32 box arithmetic: bit-exact too."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import preproc_ref as P
from tests.cases import PREPROC_CASES, preproc_boxes, preproc_image

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import letterbox, preprocess, scale_boxes

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "preproc_cases.npz")


@pytest.mark.parametrize("name", list(PREPROC_CASES))
def test_letterbox_and_preprocess_bit_exact(name):
    h, w, S, seed = PREPROC_CASES[name]
    img = preproc_image(h, w, seed)
    ref_lb, ref_ratio, ref_pad = P.letterbox(img, S)
    lb, ratio, pad = letterbox(img, S)                      # numpy in -> numpy out, like the reference
    assert isinstance(lb, np.ndarray) and lb.dtype == np.uint8
    assert np.array_equal(lb, ref_lb)
    assert hashlib.sha256(lb.tobytes()).digest() == GOLD[f"{name}/sha"].tobytes()
    assert ratio == ref_ratio and pad == ref_pad and pad == tuple(GOLD[f"{name}/pad"])
    x, ratios, pads = preprocess([torch.from_numpy(img).cuda()], S)
    assert x.shape == (1, 3, S, S) and x.dtype == torch.float32 and x.is_cuda
    xr, _, _ = P.preprocess(img, S)
    assert np.array_equal(x[0].cpu().numpy(), xr)
    assert hashlib.sha256(x[0].cpu().numpy().tobytes()).digest() == GOLD[f"{name}/x_sha"].tobytes()
    assert ratios[0] == ref_ratio and pads[0] == ref_pad


def test_preprocess_batch_pitched_source_and_custom_colour():
    imgs = [preproc_image(h, w, s) for h, w, s in [(300, 400, 11), (720, 1280, 12), (640, 480, 13)]]
    x, ratios, pads = preprocess(imgs, 640)
    for i, im in enumerate(imgs):
        xr, r, p = P.preprocess(im, 640)
        assert np.array_equal(x[i].cpu().numpy(), xr) and ratios[i] == r and pads[i] == p
    # a column crop of a wider device image: rows are not contiguous (row pitch > 3 * w)
    wide = torch.from_numpy(preproc_image(200, 500, 14)).cuda()
    crop = wide[:, 100:420]
    lb, _, _ = letterbox(crop, 640, color=(10, 20, 30))
    ref, _, _ = P.letterbox(wide.cpu().numpy()[:, 100:420], 640, color=(10, 20, 30))
    assert lb.is_cuda and np.array_equal(lb.cpu().numpy(), ref)


@pytest.mark.parametrize("name", list(PREPROC_CASES))
def test_scale_boxes_bit_exact(name):
    h, w, S, seed = PREPROC_CASES[name]
    _, ratio, pad = P.letterbox(preproc_image(h, w, seed), S)
    b = preproc_boxes(S, seed)
    out = scale_boxes(torch.from_numpy(b.copy()).cuda(), (S, S), (h, w), (ratio, pad))
    assert np.array_equal(out.cpu().numpy(), GOLD[f"{name}/boxes_rp"])
    out = scale_boxes(torch.from_numpy(b.copy()).cuda(), (S, S), (h, w))
    assert np.array_equal(out.cpu().numpy(), GOLD[f"{name}/boxes_none"])
    # in place on the xyxy columns of detection rows (row stride 6), other columns untouched
    det = torch.zeros((b.shape[0], 6), device="cuda"); det[:, :4] = torch.from_numpy(b).cuda(); det[:, 4] = 0.5; det[:, 5] = 3
    ret = scale_boxes(det[:, :4], (S, S), (h, w), (ratio, pad))
    assert ret.data_ptr() == det.data_ptr()
    assert np.array_equal(det[:, :4].cpu().numpy(), GOLD[f"{name}/boxes_rp"]) and bool((det[:, 4] == 0.5).all()) and bool((det[:, 5] == 3).all())


def test_preproc_edges():
    assert scale_boxes(torch.zeros((0, 4), device="cuda"), (640, 640), (480, 640)).shape == (0, 4)
    with pytest.raises(yolo_b200.YreError):
        scale_boxes(torch.zeros((3, 4)), (640, 640), (480, 640))                      # CPU tensor: no fallback
    with pytest.raises(yolo_b200.YreError):
        letterbox(torch.zeros((10, 10, 3), dtype=torch.uint8), 640)                   # CPU tensor: no fallback
    with pytest.raises(TypeError):
        letterbox(np.zeros((10, 10, 3), np.float32), 640)


def test_detect_script_flow_end_to_end():
    """The reference's scripts/detect.py:223-262 flow on the device: preprocess -> model -> NMS -> scale_boxes."""
    from yolo_b200 import YOLO, non_max_suppression
    from oracle import gelan_ref as G
    root = Path(__file__).resolve().parents[1]
    nodes, nc = G.load_graph(root / "configs/models/gelan-c.yaml")
    m = YOLO.from_yaml(root / "configs/models/gelan-c.yaml")
    m.load_state_dict(G.calibrated_state_dict(nodes, nc))
    m = m.cuda().eval()
    img = preproc_image(240, 320, 21)
    x, ratios, pads = preprocess(img, 320)
    y, _ = m(x)
    det = non_max_suppression(y.permute(0, 2, 1).contiguous(), 0.25, 0.45)[0]
    before = det[:, :4].clone()
    if len(det):
        det[:, :4] = scale_boxes(det[:, :4], (320, 320), img.shape[:2], (ratios[0], pads[0]))
        assert np.array_equal(det[:, :4].cpu().numpy(), P.scale_boxes(before.cpu().numpy(), (320, 320), img.shape[:2], (ratios[0], pads[0])))
        assert float(det[:, [0, 2]].max()) <= 320 and float(det[:, [1, 3]].max()) <= 240
