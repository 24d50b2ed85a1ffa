"""Checkpoint ingestion: upstream YOLOv9 / GELAN checkpoints -> this model's state_dict.

Mirrors what the reference does in two places:
  * scripts/convert_weights.py:204-249 (tables :22-95) -- upstream keys `model.<node index>.<cv1|cv2|...>` are renamed to
    the reference's `layers.<node name>.<conv_in|block1|...>`;
  * scripts/detect.py:176-182 -- a training checkpoint `{"model_state_dict": ...}` is loaded into a model built from YAML.

The renaming is not table-driven here: the node index is the position of the layer in the model graph (weight-less
nodes -- Upsample, Concat, Silence, CBFuse -- count too, exactly like the upstream YAML), and the per-block rules are
attached to this package's block classes, so any graph assembled from those blocks converts without a new table.
Pinned key-for-key against the reference's converter (tests/golden/ckpt_keys.json, tests/test_cpu_host.py).
Once loaded, the first forward folds BN / RepConv and packs bf16 weights for the B200 plan (engine.compile_model).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from pathlib import Path

import torch
from torch import nn

from . import blocks as B
from .heads import DetectDFL, DualDetectDFL

# (pattern, replacement) applied in order to the key suffix after "model.<i>."; first match of each rule only
_RULES: dict[type, list[tuple[str, str]]] = {
    B.Conv: [],
    B.CBLinear: [],
    B.ADown: [(r"^cv1\.", "conv_stride."), (r"^cv2\.", "conv_pool.")],
    B.SPPELAN: [(r"^cv1\.", "conv_in."), (r"^cv5\.", "conv_out.")],
    B.RepNCSPELAN4: [
        (r"^cv1\.", "conv_in."), (r"^cv4\.", "conv_out."),
        (r"^cv2\.0\.m\.(\d+)\.cv([12])\.", r"block1.0.bottlenecks.\1.conv\2."), (r"^cv3\.0\.m\.(\d+)\.cv([12])\.", r"block2.0.bottlenecks.\1.conv\2."),
        (r"^cv2\.0\.cv([123])\.", r"block1.0.conv\1."), (r"^cv3\.0\.cv([123])\.", r"block2.0.conv\1."),
        (r"^cv2\.", "block1."), (r"^cv3\.", "block2."),
    ],
    DetectDFL: [(r"^cv2\.", "box_convs."), (r"^cv3\.", "cls_convs.")],
    DualDetectDFL: [(r"^cv2\.", "aux_box_convs."), (r"^cv3\.", "aux_cls_convs."), (r"^cv4\.", "main_box_convs."), (r"^cv5\.", "main_cls_convs.")],
}


def _rename(suffix: str, rules: list[tuple[str, str]]) -> str:
    for pat, rep in rules:
        new, n = re.subn(pat, rep, suffix, count=1)
        if n:
            return new
    return suffix


def convert_upstream_state_dict(upstream_sd: dict, model: nn.Module) -> "OrderedDict[str, torch.Tensor]":
    """Upstream `model.<i>.*` keys -> `layers.<name>.*` keys of `model` (a yolo_b200.YOLO).  Keys that do not start with
    `model.`, and node indices without weights or beyond the graph, are skipped -- like the reference converter."""
    nodes = list(model.layers.items())
    out: OrderedDict[str, torch.Tensor] = OrderedDict()
    for key, tensor in upstream_sd.items():
        parts = key.split(".", 2)
        if len(parts) < 3 or parts[0] != "model" or not parts[1].isdigit():
            continue
        idx = int(parts[1])
        if idx >= len(nodes):
            continue
        name, mod = nodes[idx]
        rules = _RULES.get(type(mod))
        if rules is None:              # Upsample / Concat / Silence / CBFuse: no weights
            continue
        out[f"layers.{name}.{_rename(parts[2], rules)}"] = tensor
    return out


def extract_state_dict(ckpt) -> tuple[dict, str]:
    """(state_dict, kind) from anything torch.load returned: kind is 'reference' (this package's / the reference's key
    layout) or 'upstream' (yolov9 `model.<i>.` layout)."""
    if isinstance(ckpt, nn.Module):
        ckpt = {"model": ckpt}
    if not isinstance(ckpt, dict):
        raise ValueError(f"cannot extract a state_dict from {type(ckpt).__name__}")
    if "model_state_dict" in ckpt:                                   # scripts/detect.py:176-182, train/trainer checkpoints
        return ckpt["model_state_dict"], "reference"
    if "model" in ckpt and not isinstance(ckpt["model"], torch.Tensor):
        m = ckpt["model"]                                            # scripts/convert_weights.py:252-267
        sd = m.float().state_dict() if hasattr(m, "state_dict") else m
        if not isinstance(sd, dict):
            raise ValueError("checkpoint['model'] is neither a module nor a state_dict")
        ckpt = sd
    keys = [k for k in ckpt if isinstance(k, str)]
    if any(k.startswith("layers.") for k in keys):
        return ckpt, "reference"
    if any(k.startswith("model.") for k in keys):
        return ckpt, "upstream"
    raise ValueError("unrecognised checkpoint layout (expected 'layers.*', 'model.<i>.*', 'model_state_dict' or 'model')")


def load_checkpoint(model: nn.Module, ckpt, strict: bool = True) -> nn.Module:
    """Loads an upstream yolov9 / gelan checkpoint, a reference training checkpoint or a plain state_dict (object or path)
    into `model`; the compiled B200 plans are invalidated by the load hook and rebuilt on the next forward."""
    if isinstance(ckpt, (str, Path)):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd, kind = extract_state_dict(ckpt)
    if kind == "upstream":
        sd = convert_upstream_state_dict(sd, model)
    model.load_state_dict(sd, strict=strict)
    return model
