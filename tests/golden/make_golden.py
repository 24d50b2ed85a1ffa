"""Regenerates tests/golden/*.npz from the REAL reference (run in the build container, where
/root/reference is mounted; the GPU box only ever reads the committed .npz files).

    python tests/golden/make_golden.py

Weights come from oracle.gelan_ref.calibrated_state_dict (deterministic, seeds below) loaded
into the reference model with strict=True; inputs from oracle.gelan_ref.fractal.  Everything
stored here is an OUTPUT OF THE REFERENCE's own forward / non_max_suppression.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.modules.setdefault("albumentations", types.ModuleType("albumentations"))  # yolo/data/transforms.py:10
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))

from yolo import YOLO, non_max_suppression  # noqa: E402  (the reference)
from oracle import gelan_ref as G           # noqa: E402
from tests.cases import NMS_CASES, make_pred  # noqa: E402

OUT = Path(__file__).resolve().parent
torch.set_num_threads(8)


def net_case(cfg: str, size: int, batch: int, seed: int, stride_a: int):
    nodes, nc = G.load_graph(ROOT / "configs/models" / f"{cfg}.yaml")
    sd = G.calibrated_state_dict(nodes, nc)
    m = YOLO.from_yaml(f"/root/reference/configs/models/{cfg}.yaml")
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = G.fractal(batch, size, torch.Generator().manual_seed(seed))
    acts = {}
    hooks = [l.register_forward_hook(lambda _m, _i, o, n=n: acts.__setitem__(n, o) if isinstance(o, torch.Tensor) else None)
             for n, l in m.layers.items()]
    with torch.no_grad():
        y, raws = m(x)
        m64 = __import__("copy").deepcopy(m).double()
        m64.layers["detect"]._shape = None
        y64, raws64 = m64(x.double())
    for h in hooks:
        h.remove()
    if isinstance(y, list):       # dual head: callers use the main half (scripts/detect.py:239-241)
        y, raws, y64, raws64 = y[1], raws[1], y64[1], raws64[1]
    pred = y.permute(0, 2, 1).contiguous()
    dets = non_max_suppression(pred, 0.25, 0.45)
    d = {
        "cfg": cfg, "size": size, "batch": batch, "seed": seed, "stride_a": stride_a,
        "y_sub": y[:, :, ::stride_a].numpy(), "y64_sub": y64[:, :, ::stride_a].numpy(),
        "layer_names": np.array(list(acts)),
        "layer_absmean": np.array([acts[n].abs().mean().item() for n in acts], np.float64),
        "layer_mean": np.array([acts[n].mean().item() for n in acts], np.float64),
        "layer_probe": np.stack([acts[n].flatten()[:: max(1, acts[n].numel() // 64)][:64].numpy() for n in acts]),
        "floor_box": (y[:, :4] - y64[:, :4]).abs().max().item(),
        "floor_score": (y[:, 4:] - y64[:, 4:]).abs().max().item(),
        "n_cand": np.array([(pred[i, :, 4:].max(1).values > 0.25).sum().item() for i in range(batch)]),
    }
    for i, r in enumerate(raws):
        d[f"raw{i}_sub"] = r[:, :, :: max(1, stride_a // 2), :: max(1, stride_a // 2)].numpy()
    for i, t in enumerate(dets):
        d[f"det{i}"] = t.numpy()
    np.savez_compressed(OUT / f"{cfg}_{size}.npz", **d)
    print(cfg, size, "floor", d["floor_box"], d["floor_score"], "cand", d["n_cand"], "dets", [len(t) for t in dets])


def det_case(cfg: str, size: int, batch: int, seed: int):
    """Final detections (forward + non_max_suppression, conf .25 / iou .45) of the REFERENCE in fp32 on a batch of
    calibrated-weight images: the yardstick of the bf16 detection-level gate (tests/test_gpu_round2.py)."""
    nodes, nc = G.load_graph(ROOT / "configs/models" / f"{cfg}.yaml")
    sd = G.calibrated_state_dict(nodes, nc)
    m = YOLO.from_yaml(f"/root/reference/configs/models/{cfg}.yaml")
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = G.fractal(batch, size, torch.Generator().manual_seed(seed))
    with torch.no_grad():
        y, _ = m(x)
    if isinstance(y, list):
        y = y[1]
    pred = y.permute(0, 2, 1).contiguous()
    dets = non_max_suppression(pred, 0.25, 0.45)
    d = {"cfg": cfg, "size": size, "batch": batch, "seed": seed,
         "n_cand": np.array([(pred[i, :, 4:].max(1).values > 0.25).sum().item() for i in range(batch)])}
    for i, t in enumerate(dets):
        d[f"det{i}"] = t.numpy()
    np.savez_compressed(OUT / f"{cfg}_{size}_dets.npz", **d)
    print(cfg, size, "dets", [len(t) for t in dets], "cand", d["n_cand"])


def nms_cases():
    d = {}
    for name, c in NMS_CASES.items():
        p = make_pred(c)
        dets = non_max_suppression(p, **c["kw"])
        for i, t in enumerate(dets):
            d[f"{name}.{i}"] = t.numpy()
        print(name, [len(t) for t in dets])
    np.savez_compressed(OUT / "nms_cases.npz", **d)


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "nms" in only:
        nms_cases()
    if not only or "nets" in only:
        net_case("gelan-c", 128, 2, 11, 1)
        net_case("gelan-c", 640, 1, 12, 8)
        net_case("yolov9-c", 64, 1, 13, 1)
    if not only or "round2" in only:
        net_case("yolov9-c", 640, 1, 14, 8)
        det_case("gelan-c", 640, 8, 31)
