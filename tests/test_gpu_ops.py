"""GPU parity tests, stage-wise and teacher-forced: every kernel / fused block is fed the ORACLE's
input and compared with the oracle's output (SURVEY.md 8d gate 1).  Everything goes through the C
ABI (libyre.so).  fp32 validation engine: <= 1e-4 * max|ref|.  bf16 tcgen05 engine: tolerance stated
per stage below (inputs, weights and every layer output are rounded to bf16; accumulation is fp32)."""
import numpy as np
import pytest
import torch

from oracle import gelan_ref as G
from oracle import nms_ref as N
from tests.cases import NMS_CASES, make_pred

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import blocks as B, engine, _lib as L
from yolo_b200.heads import DetectDFL

DEV = "cuda"
FP32_TOL = 1e-4          # x max|ref|
BF16_CONV_TOL = 1.0e-2   # single conv, x max|ref|  (bf16 eps = 2^-8 = 3.9e-3 on in, w and out)
BF16_BLOCK_TOL = 4.0e-2  # a 12-conv ELAN block, x max|ref|


def randomize_bn(m, g):
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.8, 1.2, generator=g); mod.bias.data.normal_(0, 0.2, generator=g)
            mod.running_mean.normal_(0, 0.3, generator=g); mod.running_var.uniform_(0.5, 1.5, generator=g)
    return m.eval()


def oracle_of(m, fn, x, **kw):
    sd = {f"m.{k}": v for k, v in m.state_dict().items()}
    return getattr(G._Ctx(sd), fn)("m", x, **kw)


def run(m, x, prec):
    with yolo_b200.precision(prec):
        return m.to(DEV)(x.to(DEV)).cpu()


def check(out, ref, tol, what):
    assert out.shape == ref.shape, what
    err = (out - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"{what}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})"
    return err


# (ctor args, oracle fn, oracle kwargs, input shape) -- channel configs of gelan-c (reference tests/test_blocks.py)
CONVS = [
    ((64, 128, 3, 2), dict(stride=2), (2, 64, 32, 32)),        # stem2-like 3x3 s2 (NHWC input -> FFMA in bf16 mode)
    ((128, 64, 1, 1), {}, (2, 128, 40, 40)),
    ((64, 64, 3, 1), {}, (2, 64, 40, 40)),
    ((32, 32, 3, 1), {}, (1, 32, 24, 24)),                     # BLOCK_K = 32 / SWIZZLE_64B path
    ((64, 32, 1, 1), {}, (1, 64, 24, 24)),
    ((256, 256, 3, 1), {}, (1, 256, 20, 20)),
    ((512, 256, 1, 1), {}, (3, 512, 13, 13)),                  # ragged: 13x13, batch 3
    ((1024, 512, 1, 1), {}, (1, 1024, 20, 20)),
    ((256, 80, 1, 1), {}, (1, 256, 16, 16)),                   # N = 80
    ((64, 64, 3, 1), {}, (3, 64, 37, 29)),                     # halo kernel (8x16 patches), ragged in x and y
    ((32, 32, 3, 1), {}, (2, 32, 50, 19)),                     # halo kernel, SWIZZLE_64B, one partial patch row
    ((64, 32, 3, 1), {}, (1, 64, 48, 40)),                     # halo kernel, Cout != Cin
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("args,okw,shape", CONVS)
def test_conv(args, okw, shape, prec):
    g = torch.Generator().manual_seed(1)
    m = randomize_bn(B.Conv(*args), g)
    x = torch.randn(shape, generator=g)
    ref = oracle_of(m, "cba", x, **okw)
    check(run(m, x, prec), ref, FP32_TOL if prec == "fp32" else BF16_CONV_TOL, f"Conv{args}")


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_stem_conv_from_nchw_image(prec):
    g = torch.Generator().manual_seed(2)
    m = randomize_bn(B.Conv(3, 64, 3, 2), g)
    x = torch.rand((2, 3, 64, 96), generator=g)
    ref = oracle_of(m, "cba", x, stride=2)
    check(run(m, x, prec), ref, FP32_TOL if prec == "fp32" else BF16_CONV_TOL, "stem")


BLOCKS = [
    (lambda: B.RepConv(64, 64), "repconv", (1, 64, 20, 20), BF16_CONV_TOL),
    (lambda: B.RepNBottleneck(64, 64, expansion_ratio=1.0), "bottleneck", (1, 64, 20, 20), BF16_BLOCK_TOL),
    (lambda: B.RepNCSP(128, 128, 1), "csp", (1, 128, 20, 20), BF16_BLOCK_TOL),
    (lambda: B.RepNCSP(64, 64, 2), "csp", (2, 64, 12, 12), BF16_BLOCK_TOL),
    (lambda: B.RepNCSPELAN4(128, 256, 128, 64, 1), "elan", (1, 128, 32, 32), BF16_BLOCK_TOL),     # test_blocks.py:127-134
    (lambda: B.RepNCSPELAN4(512, 512, 512, 256, 1), "elan", (2, 512, 20, 20), BF16_BLOCK_TOL),
    (lambda: B.ADown(128, 256), "adown", (1, 128, 32, 32), BF16_CONV_TOL),                      # test_blocks.py:95-101
    (lambda: B.ADown(256, 256), "adown", (2, 256, 26, 26), BF16_CONV_TOL),                      # 416-input sizes
    (lambda: B.ADown(64, 64), "adown", (1, 64, 13, 13), BF16_CONV_TOL),                         # odd extent
    (lambda: B.SPPELAN(512, 512, 256), "sppelan", (1, 512, 20, 20), BF16_BLOCK_TOL),            # test_blocks.py:111-117
    (lambda: B.SPPELAN(64, 64, 32), "sppelan", (2, 64, 7, 9), BF16_BLOCK_TOL),
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("ctor,fn,shape,btol", BLOCKS)
def test_block(ctor, fn, shape, btol, prec):
    g = torch.Generator().manual_seed(3)
    m = randomize_bn(ctor(), g)
    x = torch.randn(shape, generator=g)
    ref = oracle_of(m, fn, x)
    check(run(m, x, prec), ref, FP32_TOL if prec == "fp32" else btol, type(m).__name__)


def test_upsample_concat_silence():
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn((2, 64, 10, 10), generator=g), torch.randn((2, 32, 20, 20), generator=g)
    with yolo_b200.precision("fp32"):
        up = B.Upsample().eval()(a.to(DEV)).cpu()
        assert torch.equal(up, torch.nn.functional.interpolate(a, scale_factor=2.0, mode="nearest"))
        cat = B.Concat().eval()([up.to(DEV), b.to(DEV)]).cpu()
        assert torch.equal(cat, torch.cat((up, b), 1))
        assert torch.equal(B.Silence().eval()(a.to(DEV)).cpu(), a)        # test_blocks.py:87-92


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_detect_head(prec):
    """DetectDFL eval: y[B,84,A] + 3 raws (reference tests/test_heads.py:82-124), numerics vs oracle."""
    g = torch.Generator().manual_seed(5)
    m = randomize_bn(DetectDFL(80, (64, 128, 128)), g)
    m.stride = torch.tensor([8.0, 16.0, 32.0])
    feats = [torch.randn((2, 64, 16, 16), generator=g), torch.randn((2, 128, 8, 8), generator=g),
             torch.randn((2, 128, 4, 4), generator=g)]
    sd = {f"m.{k}": v for k, v in m.state_dict().items()}
    cx = G._Ctx(sd)
    raws_ref = [cx.tower("m.box_convs", "m.cls_convs", i, f) for i, f in enumerate(feats)]
    y_ref = G.decode(raws_ref, [8.0, 16.0, 32.0], 80, sd["m.dfl.conv.weight"])
    with yolo_b200.precision(prec):
        y, raws = m.to(DEV)([f.to(DEV) for f in feats])
    assert y.shape == (2, 84, 16 * 16 + 8 * 8 + 4 * 4) and len(raws) == 3 and raws[0].shape == (2, 144, 16, 16)
    tol = FP32_TOL if prec == "fp32" else BF16_BLOCK_TOL
    for r, q in zip(raws, raws_ref):
        check(r.cpu(), q, tol, "raw logits")
    if prec == "fp32":
        assert (y[:, :4].cpu() - y_ref[:, :4]).abs().max() <= 1e-4 * 128
        assert (y[:, 4:].cpu() - y_ref[:, 4:]).abs().max() <= 1e-4


def test_decode_kernel_against_oracle():
    """K6 alone on random logits, fp32 and bf16 raw inputs, ragged level sizes."""
    lib = L.lib()
    import ctypes as C
    g = torch.Generator().manual_seed(6)
    nc, Bn = 80, 3
    shapes = [(12, 20), (6, 10), (3, 5)]
    strides = [8.0, 16.0, 32.0]
    raws = [torch.randn((Bn, 64 + nc, h, w), generator=g) * 3 for h, w in shapes]
    wd = torch.arange(16.0).view(1, 16, 1, 1)
    for dtype, tdt, tolb, tols in ((L.F32, torch.float32, 1e-4, 1e-5), (L.BF16, torch.bfloat16, 1e-4, 1e-5)):
        rr = [r.to(tdt).float() for r in raws]
        y_ref = G.decode(rr, strides, nc, wd).permute(0, 2, 1)
        dev_raw = [r.permute(0, 2, 3, 1).contiguous().to(DEV, tdt) for r in rr]
        A = sum(h * w for h, w in shapes)
        y = torch.empty((Bn, A, 4 + nc), device=DEV)
        d = L.DecodeDesc()
        for i, (t, (h, w)) in enumerate(zip(dev_raw, shapes)):
            d.raw[i] = L.View(t.data_ptr(), dtype, L.NHWC, Bn, h, w, 64 + nc, 0, 64 + nc)
            d.stride[i] = strides[i]
        d.levels, d.nc, d.y = 3, nc, y.data_ptr()
        for k in range(16):
            d.dfl_w[k] = float(k)
        L.check(lib.yre_dfl_decode_score(C.byref(d), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert (y[..., :4].cpu() - y_ref[..., :4]).abs().max() <= tolb * 32 * 20
        assert (y[..., 4:].cpu() - y_ref[..., 4:]).abs().max() <= tols


# ---- NMS: bit-exact -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(NMS_CASES))
def test_nms_bit_exact_vs_reference_fixture(name):
    gd = np.load("tests/golden/nms_cases.npz")
    c = NMS_CASES[name]
    p = make_pred(c)
    dets = yolo_b200.non_max_suppression(p.to(DEV), **c["kw"])
    out, counts, keep = yolo_b200.nms_raw(p.to(DEV), **c["kw"])
    odets, okeep = N.non_max_suppression(p, return_keep=True, **c["kw"])
    for i, d in enumerate(dets):
        assert np.array_equal(d.cpu().numpy(), gd[f"{name}.{i}"]), (name, i, len(d), len(gd[f"{name}.{i}"]))
        assert np.array_equal(keep[i, : counts[i]].cpu().numpy(), okeep[i])       # keep -> anchor indices, in order
        assert d.dtype == torch.float32 and d.shape[1] == 6


@pytest.mark.parametrize("A,nc,conf,iou,quant", [(8400, 80, 0.001, 0.6, None), (33600, 80, 0.001, 0.6, 4096),
                                                  (20000, 4, 0.05, 0.45, 64), (513, 80, 0.25, 0.45, None),
                                                  (70000, 8, 0.5, 0.5, None)])
def test_nms_bit_exact_large(A, nc, conf, iou, quant):
    """Stress sizes (33 600 anchors = 1280x1280, all candidates) against the C oracle."""
    from tests.cases import synth_pred
    p = synth_pred(B=2, A=A, nc=nc, S=1280, seed=A + nc, quant=quant, neg=True)
    dets = yolo_b200.non_max_suppression(p.to(DEV), conf, iou)
    ref = N.non_max_suppression(p, conf, iou, impl="c")
    for d, r in zip(dets, ref):
        assert np.array_equal(d.cpu().numpy(), r), (len(d), len(r))


def test_nms_api_edges():
    p = make_pred(NMS_CASES["plain"]).to(DEV)
    assert [tuple(d.shape) for d in yolo_b200.non_max_suppression(p, conf_thres=2.0)] == [(0, 6), (0, 6)]
    d1 = yolo_b200.non_max_suppression(p[:, :, :].permute(0, 2, 1).contiguous().permute(0, 2, 1), 0.25, 0.45)   # strided view input
    d2 = yolo_b200.non_max_suppression(p, 0.25, 0.45)
    assert all(torch.equal(a, b) for a, b in zip(d1, d2))
    with pytest.raises(ValueError):
        yolo_b200.non_max_suppression(p[0])


def test_conv_tcgen05_fuzz_vs_torch_fp32():
    """Seeded random conv shapes straight through the C ABI (yre_conv, tcgen05 engine): generic, weight-stationary halo,
    streamed halo and paired-halo schedules, channel windows, residuals, ragged edges -- against a plain torch fp32
    conv of the same bf16-rounded operands.  Tolerance: output rounding to bf16 (2^-8 relative) + accumulation order."""
    import ctypes as C
    import random
    import torch.nn.functional as F
    lib = L.lib()
    rnd = random.Random(5)
    for it in range(36):
        k = rnd.choice([1, 3, 3])
        Cin, Cout = rnd.choice([32, 64, 64, 128, 256]), rnd.choice([32, 64, 128, 160, 256, 320])
        extra = rnd.choice([0, 0, 32, 64])
        Ct, coff = Cin + extra, (rnd.choice([0, extra]) if extra else 0)
        if it % 3 == 0:          # enough exact 8x16 patches for the paired schedule
            k, Cout = 3, rnd.choice([128, 256])
            Cin = rnd.choice([64, 128]); Ct, coff = Cin, 0
            H, W = 16 * rnd.randint(2, 4), 8 * rnd.randint(3, 8)
            Bn = -(-300 // ((H // 16) * (W // 8)))
        elif it % 3 == 1:
            H, W, Bn = 16 * rnd.randint(1, 4), 8 * rnd.randint(1, 8), rnd.choice([1, 2, 3])
        else:
            H, W, Bn = rnd.randint(3, 50), rnd.randint(3, 50), rnd.choice([1, 2, 5])
        act, res = rnd.choice([0, 1]), rnd.choice([0, 1])
        g = torch.Generator().manual_seed(100 + it)
        x = torch.randn((Bn, H, W, Ct), generator=g).bfloat16()
        w = (torch.randn((Cout, k, k, Cin), generator=g) / (k * k * Cin) ** 0.5).bfloat16()
        bias = torch.randn((Cout,), generator=g)
        r = torch.randn((Bn, H, W, Cout), generator=g).bfloat16()
        ref = F.conv2d(x[..., coff:coff + Cin].float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=k // 2)
        if act:
            ref = F.silu(ref)
        if res:
            ref = ref + r.float().permute(0, 3, 1, 2)
        xd, wd, bd, rd = x.to(DEV), w.to(DEV), bias.to(DEV), r.to(DEV)
        y = torch.full((Bn, H, W, Cout), 7.0, dtype=torch.bfloat16, device=DEV)
        d = L.ConvDesc(L.View(xd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Ct, coff, Cin),
                       L.View(y.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cout, 0, Cout),
                       L.View(rd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cout, 0, Cout) if res else L.View(None, 0, 0, 0, 0, 0, 0, 0, 0),
                       wd.data_ptr(), bd.data_ptr(), k, 1, act, L.ENGINE_TCGEN05)
        L.check(lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream), "yre_conv")
        got = y.float().cpu().permute(0, 3, 1, 2)
        err = (got - ref).abs().max().item()
        assert err <= BF16_CONV_TOL * max(1.0, ref.abs().max().item()), \
            f"case {it}: B{Bn} {H}x{W} {Cin}(+{extra}@{coff})->{Cout} k{k} act{act} res{res}: max err {err:.4f}"
