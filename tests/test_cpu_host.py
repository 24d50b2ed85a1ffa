"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/yre.h
declares, the module tree is state_dict-compatible with the reference schema, the plan compiler
wires the graph correctly (checked by replaying the recorded launch list with the TEST-ONLY CPU
interpreter in tests/cpu_plan_exec.py against the oracle), and the error behaviour."""
import re
from collections import Counter

import pytest
import torch

from oracle import gelan_ref as G
from tests import cpu_plan_exec as X
from tests.conftest import ROOT

import yolo_b200
from yolo_b200 import YOLO, YreError, _lib, engine, blocks as B
from yolo_b200.heads import DetectDFL


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "yre.h").read_text()
    declared = set(re.findall(r"\b(yre_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"libyre.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.yre_version() == 200
    assert lib.yre_nms_workspace_bytes(2, 8400) > 2 * 8400 * 8


def test_error_codes_without_gpu():
    lib = _lib.lib()
    assert lib.yre_conv(None, None) == -1 and b"null" in lib.yre_last_error()
    v = _lib.View(None, 0, 0, 1, 4, 4, 8, 0, 8)
    assert lib.yre_upsample2x(v, v, None) == -1            # null data pointer
    d = _lib.NmsDesc()
    assert lib.yre_nms_batched(d, None) == -1


@pytest.mark.parametrize("cfg,n_keys,n_params", [("gelan-c", 937, 25498752), ("yolov9-c", 1460, 51182080)])
def test_state_dict_schema(cfg, n_keys, n_params):
    m = YOLO.from_yaml(ROOT / "configs/models" / f"{cfg}.yaml")
    assert m.training                                       # reference returns the model in train mode (model.py:163)
    nodes, nc = G.load_graph(ROOT / "configs/models" / f"{cfg}.yaml")
    sch = G.param_schema(nodes, nc)
    sd = m.state_dict()
    assert list(sd) == list(sch) and len(sd) == n_keys
    for k, (shape, dt) in sch.items():
        assert tuple(sd[k].shape) == tuple(shape) and sd[k].dtype == dt, k
    assert sum(p.numel() for p in m.parameters()) == n_params
    assert m.layers["detect"].stride.tolist() == [8.0, 16.0, 32.0]
    assert m.load_state_dict(G.default_state_dict(nodes, nc), strict=True).missing_keys == []
    assert set(m.connections) == set(m.layers.keys())
    # default head prior of the reference (detect.py:111-127) survives construction
    fresh = YOLO.from_yaml(ROOT / "configs/models" / f"{cfg}.yaml").state_dict()
    k = [k for k in fresh if k.endswith("cls_convs.0.2.bias")][0]
    assert abs(fresh[k][0].item() - G.default_state_dict(nodes, nc)[k][0].item()) < 1e-6


def test_num_classes_override_and_groups():
    m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml", num_classes=20)
    det = m.layers["detect"]
    assert isinstance(det, DetectDFL) and det.num_outputs == 20 + 64
    assert det.cls_convs[0][2].out_channels == 20
    assert not hasattr(m, "optim_groups")        # training API (model.py:165-203) is out of scope: not mirrored


def test_no_cpu_fallback_and_no_train_forward():
    m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml")
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 64, 64))
    m.eval()
    with pytest.raises(YreError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(YreError):
        yolo_b200.non_max_suppression(torch.zeros(1, 10, 84))
    with pytest.raises(YreError):
        B.Conv(8, 8, 1).eval()(torch.zeros(1, 8, 4, 4))


@pytest.fixture()
def dry():
    engine._ALLOW_CPU_DRY_RUN = True
    yield
    engine._ALLOW_CPU_DRY_RUN = False


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_gelan_c_plan_wiring(gelan_c, dry, prec):
    nodes, nc, sd = gelan_c
    m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml")
    m.load_state_dict(sd)
    m.eval().set_precision(prec)
    x = G.fractal(1, 128, torch.Generator().manual_seed(4))
    cap = {}
    y_ref, raws_ref = G.forward(nodes, nc, sd, x, capture=cap)
    p = engine.compile_model(m, x)
    census = Counter(n for n, _ in p.op_table())
    # 8 ELAN blocks x 12 fused convs + 5 ADown x 2 + SPP 2 + stem2 + 3 levels x 5 head convs = 124; the two Upsample
    # layers emit nothing: their consumers' 1x1 convs read the half-resolution maps (yre_conv_desc.xu)
    assert census == {"conv_ffma": 124, "adown_prepool": 5, "stem": 1, "spp_maxpool": 1, "dfl_decode_score": 1}, census
    assert p.num_launches == 132
    assert sum(1 for k, a in p.trace if k == "conv" and a.get("xu") is not None) == 2
    gf = sum(f for _, f in p.op_table()) / 1e9
    assert abs(gf - 102.136 * (128 / 640) ** 2 * 1.0) / gf < 0.06      # folded-graph FLOPs (+ dense-expanded head groups)
    X.run(p)
    _, y, raws = p.result
    tol = 2e-4 if prec == "fp32" else 0.5
    for n, v in p.vals.items():
        if isinstance(v, engine.V):
            ref = cap[n]
            err = (X.nchw(X.read(v)) - ref).abs().max().item()
            assert err <= tol * max(1.0, ref.abs().max().item()), (n, err)
    if prec == "fp32":
        yy = y.permute(0, 2, 1)
        assert (yy[:, :4] - y_ref[:, :4]).abs().max() < 2e-2
        assert (yy[:, 4:] - y_ref[:, 4:]).abs().max() < 2e-4
        for r, q in zip(raws, raws_ref):
            assert (r.t.permute(0, 3, 1, 2) - q).abs().max() < 2e-3


def test_yolov9_c_plan_wiring(yolov9_c, dry):
    nodes, nc, sd = yolov9_c
    m = YOLO.from_yaml(ROOT / "configs/models/yolov9-c.yaml")
    m.load_state_dict(sd)
    m.eval().set_precision("fp32")
    x = G.fractal(1, 64, torch.Generator().manual_seed(5))
    y_ref, _ = G.forward(nodes, nc, sd, x)
    p = engine.compile_model(m, x)
    census = Counter(n for n, _ in p.op_table())
    assert census["cbfuse_sum"] == 3 and census["stem"] == 2 and census["dfl_decode_score"] == 2
    X.run(p)
    kind, y, _ = p.result
    assert kind == "dual"
    for got, ref in zip(y, y_ref):
        yy = got.permute(0, 2, 1)
        assert (yy[:, :4] - ref[:, :4]).abs().max() < 2e-2 and (yy[:, 4:] - ref[:, 4:]).abs().max() < 2e-4


def test_yolov9_c_main_only_plan_prunes_the_aux_branch(yolov9_c, dry):
    """opt-in main_only compiles only what the main towers need (callers drop the aux half,
    scripts/detect.py:239-241); default stays API-faithful (dual)."""
    nodes, nc, sd = yolov9_c
    m = YOLO.from_yaml(ROOT / "configs/models/yolov9-c.yaml")
    m.load_state_dict(sd)
    m.eval().set_precision("fp32")
    x = G.fractal(1, 64, torch.Generator().manual_seed(5))
    (_, ym_ref), _ = G.forward(nodes, nc, sd, x)
    full = engine.compile_model(m, x)
    m.main_only = True
    p = engine.compile_model(m, x)
    census = Counter(n for n, _ in p.op_table())
    assert census.get("cbfuse_sum", 0) == 0 and census["stem"] == 1 and census["dfl_decode_score"] == 1
    gf_full, gf_main = sum(f for _, f in full.op_table()), sum(f for _, f in p.op_table())
    assert 0.40 < gf_main / gf_full < 0.46          # 102.1 of 237.6 GF: the main branch is gelan-c sized
    X.run(p)
    kind, y, raws = p.result
    assert kind == "single" and len(raws) == 3
    yy = y.permute(0, 2, 1)
    assert (yy[:, :4] - ym_ref[:, :4]).abs().max() < 2e-2 and (yy[:, 4:] - ym_ref[:, 4:]).abs().max() < 2e-4


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_block_plans(dry, prec):
    """Stand-alone block calls (what the stage-wise GPU parity tests use) wire correctly."""
    g = torch.Generator().manual_seed(0)
    cases = [(B.RepNCSPELAN4(128, 256, 128, 64, 1), "elan", (1, 128, 16, 16)),
             (B.ADown(128, 256), "adown", (1, 128, 16, 16)), (B.ADown(64, 64), "adown", (1, 64, 13, 13)),
             (B.SPPELAN(64, 64, 32), "sppelan", (1, 64, 10, 10)), (B.RepNCSP(64, 64, 2), "csp", (1, 64, 8, 8))]
    for m, fn, shape in cases:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.uniform_(0.8, 1.2, generator=g); mod.bias.data.normal_(0, 0.2, generator=g)
                mod.running_mean.normal_(0, 0.3, generator=g); mod.running_var.uniform_(0.5, 1.5, generator=g)
        m.eval()
        x = torch.randn(shape, generator=g)
        sd = {f"m.{k}": v for k, v in m.state_dict().items()}
        ref = getattr(G._Ctx(sd), fn)("m", x)
        p, out = engine.compile_module(m, x, prec)
        X.run(p)
        tol = 1e-4 if prec == "fp32" else 6e-2
        assert out.shape == ref.shape
        assert (out - ref).abs().max() <= tol * max(1.0, ref.abs().max().item()), (type(m).__name__, (out - ref).abs().max())


# ---- checkpoint ingestion --------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["gelan-c", "yolov9-c"])
def test_upstream_checkpoint_keys_match_reference_converter(cfg):
    """Key-for-key equality with the reference's scripts/convert_weights.py on the same upstream key list
    (fixture produced by the reference function, tests/golden/make_golden_ckpt.py)."""
    import json
    from yolo_b200 import convert_upstream_state_dict
    pairs = json.loads((ROOT / "tests/golden/ckpt_keys.json").read_text())[cfg]
    m = YOLO.from_yaml(ROOT / f"configs/models/{cfg}.yaml")
    upstream = {u: i for i, (u, _) in enumerate(pairs)}
    upstream["model.999.foo"] = -1           # index beyond the graph: skipped
    upstream["optimizer.state"] = -2         # not a model key: skipped
    weightless = next(i for i, mod in enumerate(m.layers.values()) if not list(mod.state_dict()))
    upstream[f"model.{weightless}.whatever"] = -3       # weight-less node (Upsample / Concat / Silence ...): skipped
    got = convert_upstream_state_dict(upstream, m)
    assert list(got.keys()) == [r for _, r in pairs]
    assert list(got.values()) == list(range(len(pairs)))
    assert set(got.keys()) == set(m.state_dict().keys())


def test_load_checkpoint_layouts(gelan_c, tmp_path):
    """Upstream layout (dict or {'model': dict}), reference training checkpoint and plain state_dict all load strictly."""
    import json
    from yolo_b200 import load_checkpoint
    nodes, nc, sd = gelan_c
    pairs = json.loads((ROOT / "tests/golden/ckpt_keys.json").read_text())["gelan-c"]
    upstream = {u: sd[r] for u, r in pairs}
    for ck in (upstream, {"model": upstream, "epoch": 3}, {"model_state_dict": sd}, sd):
        m = load_checkpoint(YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml"), ck)
        back = m.state_dict()
        assert all(torch.equal(back[k], sd[k]) for k in sd)
    path = tmp_path / "up.pt"
    torch.save({"model": upstream}, path)
    m = load_checkpoint(YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml"), path)
    assert torch.equal(m.state_dict()["layers.detect.dfl.conv.weight"], sd["layers.detect.dfl.conv.weight"])
    with pytest.raises(ValueError):
        load_checkpoint(m, {"foo": torch.zeros(1)})
    bad = dict(upstream); bad.pop(pairs[0][0])
    with pytest.raises(RuntimeError):
        load_checkpoint(YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml"), bad)           # strict: a missing key is an error


def test_ctypes_structs_match_the_c_header(tmp_path):
    """ABI drift guard: sizeof / offsetof of every descriptor in include/yre.h (compiled with gcc) equal the ctypes mirrors."""
    import ctypes as C
    import subprocess
    from yolo_b200 import _lib as L
    structs = {"yre_view": (L.View, ["ptr", "dtype", "c_off", "C"]),
               "yre_conv_desc": (L.ConvDesc, None), "yre_stem_desc": (L.StemDesc, ["x_nchw", "y", "w", "stride", "x_u8_hwc"]), "yre_decode_desc": (L.DecodeDesc, None),
               "yre_nms_desc": (L.NmsDesc, ["pred", "conf_thres", "iou_thres", "max_det", "classes", "out", "workspace_bytes", "scale"]),
               "yre_letterbox_desc": (L.LetterboxDesc, ["src", "row_pitch", "new_shape", "top", "color", "out_mode", "dst"]),
               "yre_match_desc": (L.MatchDesc, ["det", "det_stride", "gt_off", "B", "thr", "max_gt_per_image", "tp"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "yre.h"}"', "int main(void) {"]
    for name, (ct, fields) in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fields or []:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-o", str(tmp_path / "abi"), str(src)], check=True)
    out = subprocess.run([str(tmp_path / "abi")], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for name, (ct, fields) in structs.items():
        assert int(got[name]) == C.sizeof(ct), name
        for f in fields or []:
            assert int(got[f"{name}.{f}"]) == getattr(ct, f).offset, f"{name}.{f}"


@pytest.mark.parametrize("nc", [1, 3, 5])
def test_any_class_count_plan_wiring(dry, nc):
    """YOLO.from_yaml(num_classes=...) accepts any class count (reference parser.py); the plan pads the class logits to a
    multiple of 16 channels (zero weights) and the outputs expose exactly nc of them."""
    nodes, _ = G.load_graph(ROOT / "configs/models/gelan-c.yaml", num_classes=nc)
    sd = G.default_state_dict(nodes, nc, seed=nc)
    m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml", num_classes=nc)
    m.load_state_dict(sd, strict=True)
    m.eval().set_precision("fp32")
    x = torch.rand((1, 3, 64, 64), generator=torch.Generator().manual_seed(nc))
    y_ref, raws_ref = G.forward(nodes, nc, sd, x)
    p = engine.compile_model(m, x)
    X.run(p)
    _, y, raws = p.result
    yy = y.permute(0, 2, 1)
    assert yy.shape == y_ref.shape == (1, 4 + nc, 84)
    assert (yy[:, :4] - y_ref[:, :4]).abs().max() < 2e-2 and (yy[:, 4:] - y_ref[:, 4:]).abs().max() < 2e-4
    for r, q in zip(raws, raws_ref):
        got = engine._raw_nchw(r)
        assert got.shape == q.shape and (got - q).abs().max() < 2e-3


def test_uint8_to_unit_interval_is_exact_in_bf16():
    """k_stem.cu (uint8 stem) builds bf16(float(v) * (1/255)); the fp32-tensor path sees bf16(float(v) / 255)
    (scripts/detect.py:226).  The two agree for every byte value, so the two stems are bit-identical in bf16 mode."""
    v = torch.arange(256, dtype=torch.float32)
    a = (v / 255.0).bfloat16()
    b = (v * torch.tensor(1.0 / 255.0, dtype=torch.float32)).bfloat16()
    assert torch.equal(a, b)


def test_bench_inputs_match_the_oracle_recipe():
    import sys
    sys.path.insert(0, str(ROOT))
    from bench_data import octave_images
    a = octave_images(2, 96, torch.Generator().manual_seed(3))
    b = G.fractal(2, 96, torch.Generator().manual_seed(3))
    assert torch.equal(a, b)
