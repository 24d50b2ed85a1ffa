"""GPU parity tests added in round 2: the CTA-pair (cta_group::2) conv schedule, the uint8-frame stem, scale_boxes fused
into the NMS output, checkpoint ingestion on hardware, the bf16 detection-level gate, the "every conv on tcgen05"
assertion, and multi-device / multi-stream robustness.  Everything goes through the C ABI (libyre.so)."""
import ctypes as C
import json
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gelan_ref as G
from oracle import nms_ref as N
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import YOLO, _lib as L
from yolo_b200 import blocks as B

DEV = "cuda"
GOLD = ROOT / "tests" / "golden"
BF16_CONV_TOL = 1.0e-2       # x max|ref|: bf16 rounding of the output (2^-8) + accumulation order


def build(cfg, sd, prec):
    m = YOLO.from_yaml(ROOT / "configs/models" / f"{cfg}.yaml")
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval().set_precision(prec)


def _null():
    return L.View(None, 0, 0, 0, 0, 0, 0, 0, 0)


# ---- CTA pairs ----------------------------------------------------------------------------------------------------------
# (B, H, W, Cin, Cout, k, act, res, f32out): every case has an N tile >= 128 and is not halo-eligible, so the generic kernel
# runs it as clusters of two CTAs; odd / even M-tile counts, one and many rounds over the 74 clusters, both swizzle widths,
# N tiles of 128 / 160 / 256, K from 32 to 9216.
PAIR_CASES = [
    (1, 20, 20, 256, 256, 1, 1, 0, 0),       # 4 M tiles -> 2 units
    (3, 13, 13, 512, 256, 1, 1, 0, 0),       # ragged 13x13 x 3 images: odd tile out
    (1, 20, 20, 1024, 512, 1, 1, 0, 0),      # two N tiles of 256
    (16, 40, 40, 256, 256, 1, 1, 1, 0),      # 200 M tiles -> 100 units > 74 clusters, residual
    (2, 20, 20, 256, 256, 3, 1, 0, 0),       # 3x3 on a 20x20 map (no halo path), K = 2304
    (1, 40, 40, 512, 320, 3, 1, 0, 0),       # merged head conv: N tile 160
    (5, 9, 11, 64, 128, 1, 0, 0, 0),         # N tile 128, tiny ragged map
    (2, 24, 24, 32, 128, 3, 1, 0, 0),        # BLOCK_K = 32 / SWIZZLE_64B
    (7, 20, 20, 128, 128, 3, 1, 1, 0),       # odd number of M tiles (22), residual
    (2, 16, 16, 256, 160, 1, 0, 0, 1),       # fp32 output through direct stores (raw head logits)
    (1, 12, 12, 1024, 256, 3, 1, 0, 0),      # K = 9216: 144 k-iterations per tile
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_cta_pair_vs_torch_fp32(case):
    Bn, H, W, Cin, Cout, k, act, res, f32out = case
    lib = L.lib()
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn((Bn, H, W, Cin), generator=g).bfloat16()
    w = (torch.randn((Cout, k, k, Cin), generator=g) / (k * k * Cin) ** 0.5).bfloat16()
    bias = torch.randn((Cout,), generator=g)
    r = torch.randn((Bn, H, W, Cout), generator=g).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=k // 2)
    if act:
        ref = F.silu(ref)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    xd, wd, bd, rd = x.to(DEV), w.to(DEV), bias.to(DEV), r.to(DEV)
    ydt, ytt = (L.F32, torch.float32) if f32out else (L.BF16, torch.bfloat16)
    y = torch.full((Bn, H, W, Cout), 7.0, dtype=ytt, device=DEV)
    d = L.ConvDesc(L.View(xd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cin, 0, Cin),
                   L.View(y.data_ptr(), ydt, L.NHWC, Bn, H, W, Cout, 0, Cout),
                   L.View(rd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cout, 0, Cout) if res else _null(),
                   wd.data_ptr(), bd.data_ptr(), k, 1, act, L.ENGINE_TCGEN05)
    L.check(lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream), "yre_conv")
    got = y.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= BF16_CONV_TOL * max(1.0, ref.abs().max().item()), f"{case}: max err {err:.4f}"


def test_stride2_tcgen05_after_stem_teacher_forced(gelan_c):
    """stem1 (parity-plane output) -> stem2 (3x3 stride-2 on tcgen05, reading the parity planes) with the calibrated
    weights, fed the oracle's input image: the bf16 result of the pair against the oracle's fp32 stem2 activation."""
    nodes, nc, sd = gelan_c
    x = G.fractal(2, 256, torch.Generator().manual_seed(41))
    cap = {}
    G.forward(nodes, nc, sd, x, capture=cap)
    ref = cap["stem2"]
    m = build("gelan-c", sd, "bf16")
    m(x.to(DEV))
    plan = next(iter(m._plans.values()))
    names = [n for n, _ in plan.op_table()]
    assert names[0] == "stem" and names[1] == "conv_tc"          # stem2 took the tcgen05 engine, not the FFMA fallback
    v = plan.vals["stem2"]
    got = v.t[..., v.c_off:v.c_off + v.C].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    # two chained bf16 convs: twice the single-conv tolerance
    assert err <= 2 * BF16_CONV_TOL * max(1.0, ref.abs().max().item()), f"stem1->stem2 max err {err:.4f} (ref max {ref.abs().max():.2f})"


@pytest.mark.parametrize("cfg,size,batch", [("gelan-c", 320, 2), ("gelan-c", 416, 1), ("gelan-c", 640, 2), ("gelan-c", 1280, 1),
                                            ("yolov9-c", 640, 1)])
def test_every_conv_runs_on_tcgen05_in_bf16_mode(cfg, size, batch, gelan_c, yolov9_c):
    """The bf16 product mode must not silently drop to the SIMT FFMA kernel for any layer of the shipped models."""
    sd = (gelan_c if cfg == "gelan-c" else yolov9_c)[2]
    m = build(cfg, sd, "bf16")
    m(torch.zeros((batch, 3, size, size), device=DEV))
    plan = next(iter(m._plans.values()))
    table = plan.op_table()
    ffma = [(i, d) for i, ((n, _), d) in enumerate(zip(table, plan.op_descriptions())) if n == "conv_ffma"]
    assert not ffma, f"{cfg}@{size}: FFMA fallback for {ffma}"
    assert plan.num_tcgen05 == sum(1 for n, _ in table if n.startswith("conv"))


# ---- uint8 frames straight into the stem -----------------------------------------------------------
@pytest.mark.parametrize("H,W,stride,f32out", [(64, 96, 2, 0), (48, 80, 2, 0), (50, 72, 2, 0), (32, 48, 1, 0), (64, 96, 2, 1)])
def test_stem_u8_equals_fp32_tensor_path(H, W, stride, f32out):
    """yre_stem_conv fed uint8 HWC BGR frames == the same kernel fed the fp32 NCHW tensor scripts/detect.py:223-227 builds
    from them (BGR->RGB, HWC->CHW, .float() / 255), bit for bit: vector path (W % 16 == 0), generic path, fp32 output."""
    lib = L.lib()
    g = torch.Generator().manual_seed(H * W + stride)
    frames = torch.randint(0, 256, (3, H, W, 3), generator=g, dtype=torch.uint8)
    x = (frames.flip(-1).permute(0, 3, 1, 2).contiguous().float() / 255.0)
    w = torch.randn((64, 3, 3, 3), generator=g) * 0.3
    bias = torch.randn((64,), generator=g) * 0.1
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    ydt, ytt = (L.F32, torch.float32) if f32out else (L.BF16, torch.bfloat16)
    wd, bd = w.to(DEV), bias.to(DEV)
    outs = []
    for src in ("f32", "u8"):
        y = torch.zeros((3, Ho, Wo, 64), dtype=ytt, device=DEV)
        xs = (x if src == "f32" else frames).to(DEV)
        d = L.StemDesc(xs.data_ptr() if src == "f32" else None, 3, 3, H, W,
                       L.View(y.data_ptr(), ydt, L.NHWC, 3, Ho, Wo, 64, 0, 64), wd.data_ptr(), bd.data_ptr(), stride, L.ACT_SILU,
                       xs.data_ptr() if src == "u8" else None)
        L.check(lib.yre_stem_conv(C.byref(d), torch.cuda.current_stream().cuda_stream), "yre_stem_conv")
        torch.cuda.synchronize()
        outs.append(y.float().cpu())
    # the vectorised uint8 kernel shares the MMA path of the fp32-tensor kernel: bit-identical.  The generic uint8 path
    # (odd widths / stride 1, stand-alone API only -- model inputs are multiples of 32) accumulates in fp32 FMAs instead.
    if f32out or (stride == 2 and W % 16 == 0):
        assert torch.equal(outs[0], outs[1])
    ref = F.silu(F.conv2d(x, w.permute(0, 3, 1, 2).contiguous(), bias, stride=stride, padding=1)).permute(0, 2, 3, 1)
    tol = 1e-4 if f32out else BF16_CONV_TOL
    assert (outs[1] - ref).abs().max() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_model_forward_on_uint8_frames(gelan_c, prec):
    """model(uint8 [B,H,W,3] BGR frames) == model(the fp32 tensor the reference builds from them), bit for bit, and the
    detect pipeline frames -> preprocess(dtype=uint8) -> model -> NMS with fused scale_boxes equals the unfused one."""
    nodes, nc, sd = gelan_c
    g = torch.Generator().manual_seed(5)
    x = G.fractal(2, 256, g)
    frames = (x.permute(0, 2, 3, 1).flip(-1) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous()
    xf = frames.flip(-1).permute(0, 3, 1, 2).contiguous().float() / 255.0
    m = build("gelan-c", sd, prec)
    y_f, raws_f = m(xf.to(DEV))
    y_u, raws_u = m(frames.to(DEV))
    assert torch.equal(y_f, y_u) and all(torch.equal(a, b) for a, b in zip(raws_f, raws_u))
    with pytest.raises(ValueError):
        m(frames.to(DEV)[..., :2])


def test_nms_fused_scale_boxes_equals_unfused():
    from tests.cases import synth_pred
    from yolo_b200 import nms_raw, preprocess, scale_boxes, scale_rows
    p = synth_pred(B=3, A=8400, nc=80, S=640, seed=77, quant=None, neg=True).to(DEV)
    ratios = [(0.5, 0.5), (0.8333333, 0.8333333), (1.0, 1.0)]
    pads = [(0, 80), (53, 0), (0, 0)]
    origs = [(960, 1280), (768, 640), (640, 640)]
    out0, cnt0, keep0 = nms_raw(p, 0.25, 0.45)
    out1, cnt1, keep1 = nms_raw(p, 0.25, 0.45, scale=scale_rows(ratios, pads, origs, device=DEV))
    assert torch.equal(cnt0, cnt1) and torch.equal(keep0, keep1)
    for i in range(3):
        n = int(cnt0[i])
        ref = out0[i, :n].clone()
        scale_boxes(ref[:, :4], (640, 640), origs[i], (ratios[i], pads[i]))
        assert n > 0 and torch.equal(ref, out1[i, :n])
    # uint8 letterbox batch == the per-image letterbox
    g = torch.Generator().manual_seed(3)
    imgs = [torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).to(DEV) for h, w in ((480, 640), (333, 500))]
    xb, r, pd = preprocess(imgs, 320, dtype=torch.uint8)
    for i, im in enumerate(imgs):
        lb, ri, pi = yolo_b200.letterbox(im, 320)
        assert torch.equal(xb[i], lb) and ri == r[i] and pi == pd[i]


# ---- checkpoint ingestion on hardware ------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["upstream", "model_key", "model_state_dict", "module"])
def test_checkpoint_ingestion_forward_parity(gelan_c, layout, tmp_path):
    """An upstream-layout checkpoint (model.<i>.cv... keys, scripts/convert_weights.py:204-249) carrying the calibrated
    weights -> load_checkpoint -> plan -> fp32 / bf16 forward == the same weights loaded as a plain state_dict, and the
    fp32 result meets the oracle gate; the wrappings scripts/detect.py:176-182 accepts are covered."""
    from yolo_b200 import load_checkpoint
    nodes, nc, sd = gelan_c
    pairs = json.loads((GOLD / "ckpt_keys.json").read_text())["gelan-c"]
    upstream = {u: sd[r].clone() for u, r in pairs}
    if layout == "upstream":
        ck = upstream
    elif layout == "model_key":
        ck = tmp_path / "up.pt"
        torch.save({"model": upstream, "epoch": 7}, ck)
    elif layout == "model_state_dict":
        ck = {"model_state_dict": {k: v.clone() for k, v in sd.items()}, "epoch": 1}
    else:                                                       # {"model": nn.Module} as upstream yolov9 saves it
        holder = torch.nn.Module()
        holder.state_dict = lambda *a, **k: upstream          # type: ignore[assignment]
        ck = {"model": holder}
    x = G.fractal(1, 256, torch.Generator().manual_seed(17))
    y_ref, _ = G.forward(nodes, nc, sd, x)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    y64, _ = G.forward(nodes, nc, sd64, x.double())
    floor_box = (y_ref[:, :4].double() - y64[:, :4]).abs().max().item()
    floor_sc = (y_ref[:, 4:].double() - y64[:, 4:]).abs().max().item()
    try:
        m = load_checkpoint(YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml"), ck)
    except (TypeError, ValueError) as e:
        if layout == "module":
            pytest.skip(f"module-wrapped checkpoints are not accepted: {e}")
        raise
    m = m.to(DEV).eval()
    direct = build("gelan-c", sd, "fp32")
    for prec in ("fp32", "bf16"):
        ya, _ = m.set_precision(prec)(x.to(DEV))
        yb, _ = direct.set_precision(prec)(x.to(DEV))
        assert torch.equal(ya, yb), prec
        if prec == "fp32":
            dbox = (ya[:, :4].cpu().double() - y64[:, :4]).abs().max().item()
            dsc = (ya[:, 4:].cpu().double() - y64[:, 4:]).abs().max().item()
            assert dbox <= max(1e-4 * 256, 3 * floor_box) and dsc <= max(1e-4, 3 * floor_sc)


# ---- bf16 accuracy stated in detection terms -----------------------------------------------------------------------------
def _iou(a, b):
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-12)


def match_fraction(ref: np.ndarray, got: np.ndarray, thr: float = 0.9) -> float:
    """Fraction of `ref` detections that have a `got` detection of the same class with IoU >= thr."""
    if len(ref) == 0:
        return 1.0
    if len(got) == 0:
        return 0.0
    ok = (_iou(ref[:, :4], got[:, :4]) >= thr) & (ref[:, None, 5] == got[None, :, 5])
    return float(ok.any(1).mean())


# The stated bf16 tolerance of the product path, in detection terms (calibrated random weights, 640x640, batch 8,
# conf .25 / iou .45, a detection "matches" when the other set has a box of the same class with IoU >= 0.9):
#   * teacher-forced head: the bf16 towers + decode + NMS fed the fp32 engine's neck features reproduce at least
#     BF16_HEAD_AGREEMENT of the fp32 REFERENCE's final detections, and vice versa;
#   * end to end the calibrated random-weight network is CHAOTIC -- the reference's own fp32-vs-fp64 difference is already
#     amplified x300 from stem to pan2, so a 2^-9 rounding per layer decorrelates the deep features
#     (rel-L2 ~0.6 at pan2) for ANY bf16 implementation.  The end-to-end agreement is therefore printed next to the same
#     figure for the reference graph itself run in bf16 (model.bfloat16() semantics) and gated relative to that yardstick.
BF16_HEAD_AGREEMENT = 0.90


def test_bf16_detection_level_agreement_with_reference(gelan_c):
    nodes, nc, sd = gelan_c
    gd = np.load(GOLD / "gelan-c_640_dets.npz")                 # produced by the real reference (make_golden.py round2)
    Bn, S = int(gd["batch"]), int(gd["size"])
    x = G.fractal(Bn, S, torch.Generator().manual_seed(int(gd["seed"])))
    ref = [gd[f"det{i}"] for i in range(Bn)]

    def agreement(dets, idx=None):
        idx = range(Bn) if idx is None else idx
        rec = [match_fraction(ref[i], dets[j]) for j, i in enumerate(idx)]
        pre = [match_fraction(dets[j], ref[i]) for j, i in enumerate(idx)]
        return float(np.mean(rec)), float(np.mean(pre))

    # fp32 validation engine: reproduces the reference's detections up to its own fp32 noise floor
    m32 = build("gelan-c", sd, "fp32")
    y32, _ = m32(x.to(DEV))
    d32 = [d.cpu().numpy() for d in yolo_b200.non_max_suppression(y32.permute(0, 2, 1), 0.25, 0.45)]
    r32 = agreement(d32)
    print(f"fp32 engine: recall {r32[0]:.3f} precision {r32[1]:.3f} of the reference's {sum(len(r) for r in ref)} detections")
    assert r32[0] >= 0.97 and r32[1] >= 0.97

    # bf16 head, teacher-forced with the fp32 engine's neck features
    p32 = next(iter(m32._plans.values()))
    det_name = list(m32.layers.keys())[-1]
    feats = []
    for n in m32.connections[det_name]:
        v = p32.vals[n]
        feats.append(v.t[..., v.c_off:v.c_off + v.C].float().permute(0, 3, 1, 2).contiguous())
    with yolo_b200.precision("bf16"):
        yh, _ = m32.layers[det_name](feats)
    dh = [d.cpu().numpy() for d in yolo_b200.non_max_suppression(yh.permute(0, 2, 1), 0.25, 0.45)]
    rh = agreement(dh)
    print(f"bf16 head (teacher-forced): recall {rh[0]:.3f} precision {rh[1]:.3f}")
    assert rh[0] >= BF16_HEAD_AGREEMENT and rh[1] >= BF16_HEAD_AGREEMENT

    # bf16 end to end, next to the reference graph run in bf16 (first two images: the CPU bf16 run is slow)
    m16 = build("gelan-c", sd, "bf16")
    y16, _ = m16(x.to(DEV))
    d16 = [d.cpu().numpy() for d in yolo_b200.non_max_suppression(y16.permute(0, 2, 1), 0.25, 0.45)]
    r16 = agreement(d16)
    sdb = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd.items()}
    yb, _ = G.forward(nodes, nc, sdb, x[:2].bfloat16())
    db = N.non_max_suppression(yb.float().permute(0, 2, 1).contiguous(), 0.25, 0.45)
    rb = agreement(db, idx=[0, 1])
    r16_2 = agreement(d16[:2], idx=[0, 1])
    print(f"bf16 end to end: recall {r16[0]:.3f} precision {r16[1]:.3f} (8 images); on images 0-1: ours {r16_2[0]:.3f}/{r16_2[1]:.3f}, "
          f"reference graph in bf16 {rb[0]:.3f}/{rb[1]:.3f}")
    assert r16_2[0] >= 0.5 * rb[0] - 0.02 and r16_2[1] >= 0.5 * rb[1] - 0.02


def test_yolov9c_640_vs_reference_fixture(yolov9_c):
    """yolov9-c (Silence / CBLinear / CBFuse / DualDetectDFL) at the full 640x640, fp32 validation mode against the REAL
    reference's forward (fixture), main head, incl. its final detections."""
    nodes, nc, sd = yolov9_c
    gd = np.load(GOLD / "yolov9-c_640.npz")
    S, Bn, sa = int(gd["size"]), int(gd["batch"]), int(gd["stride_a"])
    x = G.fractal(Bn, S, torch.Generator().manual_seed(int(gd["seed"])))
    m = build("yolov9-c", sd, "fp32")
    (ya, ym), (ra, rm) = m(x.to(DEV))
    assert ya.shape == ym.shape == (Bn, 84, 8400)
    ys = ym[:, :, ::sa].cpu().numpy()
    tol_box = max(1e-4 * S, 3 * float(gd["floor_box"]))
    tol_sc = max(1e-4, 3 * float(gd["floor_score"]))
    dbox = np.abs(ys[:, :4] - gd["y64_sub"][:, :4]).max()
    dsc = np.abs(ys[:, 4:] - gd["y64_sub"][:, 4:]).max()
    print(f"yolov9-c 640: |dbox|={dbox:.3e} (tol {tol_box:.3e})  |dscore|={dsc:.3e} (tol {tol_sc:.3e})")
    assert dbox <= tol_box and dsc <= tol_sc
    dets = yolo_b200.non_max_suppression(ym.permute(0, 2, 1), 0.25, 0.45)
    rec = match_fraction(gd["det0"], dets[0].cpu().numpy())
    assert rec >= 0.97, rec


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cblinear_cbfuse_teacher_forced(prec):
    """yolov9-c's auxiliary routing, stage-wise: CBLinear (1x1 conv with bias, no BN / activation, split) and CBFuse
    (nearest resize of the selected routes to the target size + sum) on given inputs against torch fp32."""
    g = torch.Generator().manual_seed(23)
    lin_a, lin_b = B.CBLinear(64, [32]).eval(), B.CBLinear(128, [32, 64]).eval()
    for lin in (lin_a, lin_b):
        lin.conv.weight.data.normal_(0, 0.1, generator=g); lin.conv.bias.data.normal_(0, 0.2, generator=g)
    xa, xb = torch.randn((2, 64, 16, 16), generator=g), torch.randn((2, 128, 8, 8), generator=g)
    tgt = torch.randn((2, 32, 4, 4), generator=g)
    if prec == "bf16":                       # teacher-forced: both sides see the bf16-rounded inputs
        xa, xb, tgt = xa.bfloat16().float(), xb.bfloat16().float(), tgt.bfloat16().float()
    ra = F.conv2d(xa, lin_a.conv.weight, lin_a.conv.bias).split([32], 1)
    rb = F.conv2d(xb, lin_b.conv.weight, lin_b.conv.bias).split([32, 64], 1)
    ref = tgt + F.interpolate(ra[0], size=(4, 4), mode="nearest") + F.interpolate(rb[0], size=(4, 4), mode="nearest")
    tol = 1e-4 if prec == "fp32" else BF16_CONV_TOL
    with yolo_b200.precision(prec):
        oa = lin_a.to(DEV)(xa.to(DEV))
        ob = lin_b.to(DEV)(xb.to(DEV))
        for got, want in zip(list(oa) + list(ob), list(ra) + list(rb)):
            assert (got.cpu() - want).abs().max() <= tol * max(1.0, want.abs().max().item())
        fused = B.CBFuse([0, 0]).eval()([tuple(o for o in oa), tuple(o for o in ob), tgt.to(DEV)])
    # the fuse consumes the (bf16-rounded) projections: three roundings on the way
    assert (fused.cpu() - ref).abs().max() <= 3 * tol * max(1.0, ref.abs().max().item())


# ---- robustness -----------------------------------------------------------------------------------------------------------
def test_two_devices_in_one_process(gelan_c):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a process that runs on cuda:0 and then cuda:1 must
    opt in on both (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nodes, nc, sd = gelan_c
    x = G.fractal(1, 256, torch.Generator().manual_seed(3))
    ys = []
    for d in ("cuda:0", "cuda:1"):
        m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml")
        m.load_state_dict(sd, strict=True)
        m = m.to(d).eval()
        y, _ = m(x.to(d))
        dets = yolo_b200.non_max_suppression(y.permute(0, 2, 1), 0.25, 0.45)
        ys.append((y.cpu(), [t.cpu() for t in dets]))
    assert torch.equal(ys[0][0], ys[1][0]) and all(torch.equal(a, b) for a, b in zip(ys[0][1], ys[1][1]))


def test_nms_concurrent_streams_do_not_share_scratch():
    """Two NMS calls in flight on different streams (different predictions, same shape) give the results of the serial calls."""
    from tests.cases import synth_pred
    pa = synth_pred(B=4, A=8400, nc=80, S=640, seed=1, quant=None, neg=True).to(DEV)
    pb = synth_pred(B=4, A=8400, nc=80, S=640, seed=2, quant=None, neg=True).to(DEV)
    ra, rb = yolo_b200.nms_raw(pa, 0.05, 0.45), yolo_b200.nms_raw(pb, 0.05, 0.45)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        with torch.cuda.stream(s1):
            qa = yolo_b200.nms_raw(pa, 0.05, 0.45)
        with torch.cuda.stream(s2):
            qb = yolo_b200.nms_raw(pb, 0.05, 0.45)
        torch.cuda.synchronize()
        for r, q in ((ra, qa), (rb, qb)):
            assert torch.equal(r[1], q[1])
            for i in range(4):
                n = int(r[1][i])
                assert torch.equal(r[0][i, :n], q[0][i, :n]) and torch.equal(r[2][i, :n], q[2][i, :n])


def test_async_nms_and_graph_mode_with_fresh_inputs(gelan_c):
    """non_max_suppression_async().result() == non_max_suppression(); CUDA-graph mode fed a NEW input tensor every call
    stays correct and holds at most 8 (graph, buffer) pairs."""
    nodes, nc, sd = gelan_c
    m = build("gelan-c", sd, "bf16")
    m.fresh_outputs, m.use_cuda_graph = False, True
    g = torch.Generator().manual_seed(8)
    base = G.fractal(1, 128, g).to(DEV)
    for i in range(11):
        x = (base * (0.5 + 0.04 * i)).clone()
        y = m(x)[0].clone()
        m.use_cuda_graph = False
        assert torch.equal(m(x)[0], y)
        m.use_cuda_graph = True
    plan = next(iter(m._plans.values()))
    assert len(plan.graphs) <= 8
    pred = y.permute(0, 2, 1)
    a = yolo_b200.non_max_suppression_async(pred, 0.25, 0.45).result()
    b = yolo_b200.non_max_suppression(pred, 0.25, 0.45)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


@pytest.mark.parametrize("nc", [1, 3, 5, 80])
def test_any_class_count_decodes(nc):
    """YOLO.from_yaml(num_classes=...) accepts any class count (reference parser); the decode kernel must too."""
    m = YOLO.from_yaml(ROOT / "configs/models/gelan-c.yaml", num_classes=nc)
    nodes, _ = G.load_graph(ROOT / "configs/models/gelan-c.yaml", num_classes=nc)
    sd = G.default_state_dict(nodes, nc, seed=nc)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval().set_precision("fp32")
    x = torch.rand((1, 3, 64, 64), generator=torch.Generator().manual_seed(nc))
    y, raws = m(x.to(DEV))
    y_ref, _ = G.forward(nodes, nc, sd, x)
    assert y.shape == (1, 4 + nc, 84)
    assert (y[:, :4].cpu() - y_ref[:, :4]).abs().max() <= 1e-2 and (y[:, 4:].cpu() - y_ref[:, 4:]).abs().max() <= 1e-4
    dets = yolo_b200.non_max_suppression(y.permute(0, 2, 1), 0.0001, 0.45)
    ref = N.non_max_suppression(y.permute(0, 2, 1).contiguous().cpu(), 0.0001, 0.45)
    assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(dets, ref))
