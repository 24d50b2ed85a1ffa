"""GPU parity of the whole path through the public API (YOLO.from_yaml / forward /
non_max_suppression) against the oracle and the reference-generated fixtures."""
import numpy as np
import pytest
import torch

from oracle import gelan_ref as G
from oracle import nms_ref as N
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import YOLO

DEV = "cuda"
GOLD = ROOT / "tests" / "golden"


def build(cfg, sd, prec):
    m = YOLO.from_yaml(ROOT / "configs/models" / f"{cfg}.yaml")
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval().set_precision(prec)


@pytest.mark.parametrize("name,fix", [("gelan-c_128", "gelan_c"), ("gelan-c_640", "gelan_c"), ("yolov9-c_64", "yolov9_c")])
def test_fp32_end_to_end_vs_reference_fixture(name, fix, request):
    """fp32 validation mode vs the REFERENCE's own forward (fixture): |dbox| <= 1e-4*S px and
    |dscore| <= 1e-4, or 3x the reference's own fp32-vs-fp64 floor where that floor is larger
    (SURVEY.md 8d gate 2); detections of our NMS on our y equal the reference's on its y where the
    candidate sets are unambiguous."""
    nodes, nc, sd = request.getfixturevalue(fix)
    gd = np.load(GOLD / f"{name}.npz")
    cfg, S, Bn, sa = str(gd["cfg"]), int(gd["size"]), int(gd["batch"]), int(gd["stride_a"])
    x = G.fractal(Bn, S, torch.Generator().manual_seed(int(gd["seed"])))
    m = build(cfg, sd, "fp32")
    y, raws = m(x.to(DEV))
    if isinstance(y, list):
        y, raws = y[1], raws[1]
    assert y.shape[1] == 84                                   # reference tests/test_model.py:62
    ys = y[:, :, ::sa].cpu().numpy()
    tol_box = max(1e-4 * S, 3 * float(gd["floor_box"]))
    tol_sc = max(1e-4, 3 * float(gd["floor_score"]))
    dbox = np.abs(ys[:, :4] - gd["y64_sub"][:, :4]).max()
    dsc = np.abs(ys[:, 4:] - gd["y64_sub"][:, 4:]).max()
    print(f"{name}: |dbox|={dbox:.3e} (tol {tol_box:.3e}, ref floor {float(gd['floor_box']):.3e})  "
          f"|dscore|={dsc:.3e} (tol {tol_sc:.3e}, ref floor {float(gd['floor_score']):.3e})")
    assert dbox <= tol_box and dsc <= tol_sc
    for i, r in enumerate(raws):
        s = max(1, sa // 2)
        assert np.abs(r[:, :, ::s, ::s].cpu().numpy() - gd[f"raw{i}_sub"]).max() <= 2e-3


@pytest.mark.parametrize("size,batch", [(320, 1), (416, 2), (640, 1), (256, 4), (1280, 1)])
def test_fp32_end_to_end_vs_oracle_and_nms(gelan_c, size, batch):
    """Input sizes 320/416/640 and batch 1/2/4 (reference tests/test_model.py:80-95), numerics vs the
    oracle, then NMS through the public API bit-exact vs the oracle NMS on the SAME predictions."""
    nodes, nc, sd = gelan_c
    x = G.fractal(batch, size, torch.Generator().manual_seed(size))
    y_ref, raws_ref = G.forward(nodes, nc, sd, x)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    y64, _ = G.forward(nodes, nc, sd64, x.double())               # the reference graph in fp64 = ground truth
    floor_box = (y_ref[:, :4].double() - y64[:, :4]).abs().max().item()      # the reference's own fp32 noise
    floor_sc = (y_ref[:, 4:].double() - y64[:, 4:]).abs().max().item()
    m = build("gelan-c", sd, "fp32")
    y, raws = m(x.to(DEV))
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    assert y.shape == (batch, 84, A) and [tuple(r.shape) for r in raws] == [(batch, 144, size // s, size // s) for s in (8, 16, 32)]
    dbox = (y[:, :4].cpu().double() - y64[:, :4]).abs().max().item()
    dsc = (y[:, 4:].cpu().double() - y64[:, 4:]).abs().max().item()
    print(f"{size}x{size} B{batch}: ours-vs-fp64 |dbox|={dbox:.3e}px |dscore|={dsc:.3e}; reference fp32-vs-fp64 floor {floor_box:.3e}px {floor_sc:.3e}")
    assert dbox <= max(1e-4 * size, 3 * floor_box) and dsc <= max(1e-4, 3 * floor_sc)
    pred = y.permute(0, 2, 1).contiguous()                       # what callers do (scripts/detect.py:247)
    dets = yolo_b200.non_max_suppression(pred, 0.25, 0.45)
    ref = N.non_max_suppression(pred.cpu(), 0.25, 0.45)
    for a, b in zip(dets, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    # a second call returns fresh tensors (the first result is not overwritten)
    y2, _ = m((x * 0.5).to(DEV))
    assert not torch.equal(y, y2) and y.data_ptr() != y2.data_ptr()


def test_default_init_end_to_end(gelan_c):
    """Literal default init (what BASELINE.json's configs name): bias-dominated outputs, 1e-4 gate."""
    nodes, nc, _ = gelan_c
    sd = G.default_state_dict(nodes, nc)
    x = torch.rand((1, 3, 320, 320), generator=torch.Generator().manual_seed(7))
    y_ref, _ = G.forward(nodes, nc, sd, x)
    # fp32 validation mode meets the 1e-4 gate; bf16 is gated at its stated tolerance (box logits are
    # 1.0 + O(1e-3): the bf16 rounding of the O(1e-3) part is amplified x~680 by DFL * stride 32)
    for prec, tb, ts in (("fp32", 1e-4 * 320, 1e-4), ("bf16", 0.5, 1e-4)):
        y, _ = build("gelan-c", sd, prec)(x.to(DEV))
        assert (y[:, :4].cpu() - y_ref[:, :4]).abs().max() <= tb, prec
        assert (y[:, 4:].cpu() - y_ref[:, 4:]).abs().max() <= ts, prec


def test_bf16_end_to_end_drift_reported(gelan_c):
    """bf16 product path end to end: drift on random-weight nets is large for ANY bf16 implementation
    (the reference's own .bfloat16() run drifts more); it is reported and only loosely
    gated -- the per-stage tolerances in test_gpu_ops.py are the real gate."""
    nodes, nc, sd = gelan_c
    x = G.fractal(2, 640, torch.Generator().manual_seed(12))
    y_ref, _ = G.forward(nodes, nc, sd, x)
    m = build("gelan-c", sd, "bf16")
    y, _ = m(x.to(DEV))
    plan = next(iter(m._plans.values()))
    assert plan.num_tcgen05 >= 120, f"only {plan.num_tcgen05} convs ran on tcgen05"
    dbox = (y[:, :4].cpu() - y_ref[:, :4]).abs()
    dsc = (y[:, 4:].cpu() - y_ref[:, 4:]).abs()
    # yardstick: the reference graph itself run in bf16 (model.bfloat16() semantics: three roundings per Conv)
    sdb = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd.items()}
    yb, _ = G.forward(nodes, nc, sdb, x.bfloat16())
    rbox, rsc = (yb[:, :4].float() - y_ref[:, :4]).abs(), (yb[:, 4:].float() - y_ref[:, 4:]).abs()
    print(f"bf16 e2e drift: box mean {dbox.mean():.3f}px max {dbox.max():.2f}px; score mean {dsc.mean():.2e} max {dsc.max():.2e}; "
          f"reference-in-bf16: box mean {rbox.mean():.3f}px, score mean {rsc.mean():.2e}; "
          f"tcgen05 convs {plan.num_tcgen05}/{plan.num_launches} launches")
    assert dbox.mean() <= 1.25 * rbox.mean() + 0.05 and dsc.mean() <= 1.25 * rsc.mean() + 1e-4
    dets = yolo_b200.non_max_suppression(y.permute(0, 2, 1), 0.25, 0.45)
    ref = N.non_max_suppression(y.permute(0, 2, 1).contiguous().cpu(), 0.25, 0.45)
    for a, b in zip(dets, ref):
        assert np.array_equal(a.cpu().numpy(), b)


def test_config5_1280_eval_regime_nms_bit_exact(gelan_c):
    """BASELINE.json config 5: 1280x1280 (33 600 anchors/image), eval-time NMS thresholds conf 0.001 / iou 0.6,
    max_det 300 -- the bf16 product path end to end, then NMS bit-exact vs the oracle on the SAME predictions."""
    nodes, nc, sd = gelan_c
    x = G.fractal(2, 1280, torch.Generator().manual_seed(13))
    m = build("gelan-c", sd, "bf16")
    y, raws = m(x.to(DEV))
    assert y.shape == (2, 84, 33600) and [tuple(r.shape[2:]) for r in raws] == [(160, 160), (80, 80), (40, 40)]
    pred = y.permute(0, 2, 1).contiguous()
    n_cand = (pred[:, :, 4:].max(2).values > 0.001).sum(1)
    assert int(n_cand.min()) > 20000                                  # the stress regime: (almost) every anchor is a candidate
    dets = yolo_b200.non_max_suppression(pred, 0.001, 0.6, max_det=300)
    ref = N.non_max_suppression(pred.cpu(), 0.001, 0.6, max_det=300)
    for a, b in zip(dets, ref):
        assert a.shape[0] == 300 and np.array_equal(a.cpu().numpy(), b)


def test_yolov9c_dual_head_outputs(yolov9_c):
    nodes, nc, sd = yolov9_c
    x = G.fractal(1, 128, torch.Generator().manual_seed(9))
    (ya_ref, ym_ref), _ = G.forward(nodes, nc, sd, x)
    m = build("yolov9-c", sd, "fp32")
    (ya, ym), (ra, rm) = m(x.to(DEV))
    assert ya.shape == ym.shape == (1, 84, 336) and len(ra) == len(rm) == 3
    for got, ref in ((ya, ya_ref), (ym, ym_ref)):
        assert (got[:, :4].cpu() - ref[:, :4]).abs().max() <= 2e-2 and (got[:, 4:].cpu() - ref[:, 4:]).abs().max() <= 2e-4
    m.set_precision("bf16")
    (ya16, ym16), _ = m(x.to(DEV))
    assert (ym16[:, 4:].cpu() - ym_ref[:, 4:]).abs().mean() < 5e-3


def test_yolov9c_main_only_equals_main_half(yolov9_c):
    """Opt-in aux-branch pruning: same kernels on the same data, so the result equals the dual model's main half bit for bit."""
    nodes, nc, sd = yolov9_c
    x = G.fractal(2, 128, torch.Generator().manual_seed(10)).to(DEV)
    for prec in ("fp32", "bf16"):
        m = build("yolov9-c", sd, prec)
        (_, ym), (_, rm) = m(x)
        n_full = m._plans[next(iter(m._plans))].num_launches
        m.main_only = True
        y1, r1 = m(x)
        n_main = [p for k, p in m._plans.items() if k[-1]][0].num_launches
        assert n_main < 0.7 * n_full            # 134 of 215 launches
        assert torch.equal(y1, ym) and all(torch.equal(a, b) for a, b in zip(r1, rm))
        m.main_only = False
        (_, ym2), _ = m(x)
        assert torch.equal(ym2, ym)


def test_cuda_graph_replay_equals_eager(gelan_c):
    """Static-buffer mode + use_cuda_graph: the captured launch list (with its PDL edges) reproduces the eager result
    bit for bit, per input buffer, across replays."""
    nodes, nc, sd = gelan_c
    xa = G.fractal(2, 320, torch.Generator().manual_seed(21)).to(DEV)
    xb = G.fractal(2, 320, torch.Generator().manual_seed(22)).to(DEV)
    m = build("gelan-c", sd, "bf16")
    ya, yb = m(xa)[0].clone(), m(xb)[0].clone()
    m.fresh_outputs, m.use_cuda_graph = False, True
    for _ in range(3):
        assert torch.equal(m(xa)[0], ya)
        assert torch.equal(m(xb)[0], yb)
    xa.mul_(0.5)                                    # same buffer, new contents: the replay reads the buffer, not a copy
    y3 = m(xa)[0].clone()
    m.use_cuda_graph = False
    assert torch.equal(m(xa)[0], y3) and not torch.equal(y3, ya)


def test_fresh_outputs_rebind_fp32_tma_store(gelan_c):
    """Default mode returns NEW tensors every call: the fp32 raw-logit outputs leave through a TMA store, so re-binding
    them re-encodes the output tensor map (conv_tc_rebind).  Earlier results must stay untouched and equal."""
    nodes, nc, sd = gelan_c
    x = G.fractal(2, 320, torch.Generator().manual_seed(5)).to(DEV)
    m = build("gelan-c", sd, "bf16")
    y1, r1 = m(x)
    keep = [r.clone() for r in r1]
    y2, r2 = m(x)
    y3, r3 = m(x)
    torch.cuda.synchronize()
    assert all(a.data_ptr() != b.data_ptr() for a, b in zip(r1, r2)) and y1.data_ptr() != y2.data_ptr()
    assert torch.equal(y1, y2) and torch.equal(y2, y3)
    for a, b, c, k in zip(r1, r2, r3, keep):
        assert torch.equal(a, k) and torch.equal(b, k) and torch.equal(c, k)
        assert torch.isfinite(a).all()


def test_state_dict_roundtrip_and_replan(gelan_c):
    nodes, nc, sd = gelan_c
    m = build("gelan-c", sd, "fp32")
    x = G.fractal(1, 128, torch.Generator().manual_seed(1)).to(DEV)
    y1, _ = m(x)
    back = {k: v.cpu() for k, v in m.state_dict().items()}
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    sd2 = {k: (v * 1.5 if k.endswith("stem1.bn.weight") else v) for k, v in sd.items()}
    m.load_state_dict(sd2, strict=True)                           # must invalidate the compiled plan
    y2, _ = m(x)
    y2_ref, _ = G.forward(nodes, nc, sd2, x.cpu())
    assert not torch.allclose(y1, y2) and (y2[:, 4:].cpu() - y2_ref[:, 4:]).abs().max() <= 1e-3
    with torch.no_grad():                                         # in-place edit is detected through tensor versions
        m.layers["stem1"].bn.weight.mul_(1 / 1.5)
    y3, _ = m(x)
    assert (y3[:, 4:] - y1[:, 4:]).abs().max() <= 1e-5
