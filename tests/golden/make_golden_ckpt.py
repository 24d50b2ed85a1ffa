"""Regenerates tests/golden/ckpt_keys.json: (upstream yolov9 key -> reference key) pairs produced by the REFERENCE's own
converter (scripts/convert_weights.py:204-249, tables :22-95) for gelan-c and yolov9-c.  Run in the build container.

    python tests/golden/make_golden_ckpt.py

The upstream checkpoints themselves are not available offline, so the upstream-format key list is synthesised from the
reference model's state_dict by inverting the renaming rules documented in the converter's docstrings (cv1/cv2/... names,
`model.<node index>.` prefix); the inversion is validated here: converting the synthetic keys with the reference's
function must give back exactly the reference model's key set."""
import importlib.util
import json
import re
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.modules.setdefault("albumentations", types.ModuleType("albumentations"))
sys.path.insert(0, "/root/reference/src")
spec = importlib.util.spec_from_file_location("ref_convert", "/root/reference/scripts/convert_weights.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
from yolo import YOLO  # noqa: E402  (the reference)

INV = {  # reference block type -> ordered (reference substring, upstream substring) rewrites applied to the part after the layer name
    "Conv": [], "CBLinear": [],
    "ADown": [("conv_stride.", "cv1."), ("conv_pool.", "cv2.")],
    "SPPELAN": [("conv_in.", "cv1."), ("conv_out.", "cv5.")],
    "DetectDFL": [("box_convs.", "cv2."), ("cls_convs.", "cv3.")],
    "DualDetectDFL": [("aux_box_convs.", "cv2."), ("aux_cls_convs.", "cv3."), ("main_box_convs.", "cv4."), ("main_cls_convs.", "cv5.")],
}


def to_upstream(rest: str, typ: str) -> str:
    if typ != "RepNCSPELAN4":
        for a, b in INV[typ]:
            rest = rest.replace(a, b)
        return rest
    rest = re.sub(r"^conv_in\.", "cv1.", rest)
    rest = re.sub(r"^conv_out\.", "cv4.", rest)
    m = re.match(r"^(block[12])\.0\.(.*)$", rest)
    if m:
        inner = m.group(2)
        inner = re.sub(r"^bottlenecks\.(\d+)\.conv([12])\.", r"m.\1.cv\2.", inner)
        inner = re.sub(r"^conv([123])\.", r"cv\1.", inner)
        rest = f"{m.group(1)}.0.{inner}"
    rest = re.sub(r"^block1\.", "cv2.", rest)
    rest = re.sub(r"^block2\.", "cv3.", rest)
    return rest


out = {}
for name, table in (("gelan-c", ref.GELAN_C_LAYERS), ("yolov9-c", ref.YOLOV9_C_LAYERS)):
    model = YOLO.from_yaml(f"/root/reference/configs/models/{name}.yaml")
    keys = list(model.state_dict().keys())
    by_layer = {ln: (idx, typ) for idx, (ln, typ) in table.items()}
    pairs = []
    for k in keys:
        _, ln, rest = k.split(".", 2)
        idx, typ = by_layer[ln]
        pairs.append((f"model.{idx}.{to_upstream(rest, typ)}", k))
    fake = {u: i for i, (u, _) in enumerate(pairs)}
    fake["model.999.foo"] = -1            # index without weights: skipped by the reference
    fake["optimizer.state"] = -2          # not a model key: skipped
    conv = ref.convert_state_dict(fake, table)
    assert list(conv.keys()) == keys and list(conv.values()) == list(range(len(keys))), name
    out[name] = pairs
    print(name, len(pairs), "pairs; reference converter round-trips")
(Path(__file__).resolve().parent / "ckpt_keys.json").write_text(json.dumps(out, indent=0))
