"""GPU parity tests of the virtual Upsample + Concat input (yre_conv_desc.xu): the neck's ``Upsample -> Concat -> ELAN``
(configs/models/gelan-c.yaml up1/concat1/fpn1, up2/concat2/fpn2; src/yolo/model/parser.py:159-171, blocks/common.py:32-33,
blocks/gelan.py:58) runs without the upsampled tensor ever being written: the consumer's 1x1 conv reads the half-resolution
map through a tensor map with two stride-0 dimensions.  Everything goes through the C ABI (libyre.so)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import gelan_ref as G
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

import yolo_b200
from yolo_b200 import YOLO, _lib as L

DEV = "cuda"


def _null():
    return L.View(None, 0, 0, 0, 0, 0, 0, 0, 0)


def _conv(lib, x, y, w, b, act, engine, xu=None):
    Bn, H, W, Cx = x.shape
    Cout = y.shape[3]
    dt = L.BF16 if x.dtype == torch.bfloat16 else L.F32
    d = L.ConvDesc(L.View(x.data_ptr(), dt, L.NHWC, Bn, H, W, Cx, 0, Cx),
                   L.View(y.data_ptr(), L.BF16 if y.dtype == torch.bfloat16 else L.F32, L.NHWC, Bn, H, W, Cout, 0, Cout), _null(),
                   w.data_ptr(), b.data_ptr(), 1, 1, act, engine,
                   L.View(xu.data_ptr(), dt, L.NHWC, xu.shape[0], xu.shape[1], xu.shape[2], xu.shape[3], 0, xu.shape[3]) if xu is not None else _null())
    return lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream)


# (B, H, W, Cu, Cx, Cout): one box per image of a tile (tb = 1, 2, 8), images beyond the batch in the last tile, ragged
# maps, CTA pairs (wide N, long K), both swizzle widths (K chunks of 64 and of 32), one or many rounds over the SMs
UP_CASES = [
    (4, 40, 40, 512, 512, 512),      # gelan-c fpn1.conv_in at 640x640: 8x8 patches of two images per tile, CTA pairs
    (3, 80, 80, 512, 512, 256),      # gelan-c fpn2.conv_in: one image per tile
    (5, 20, 20, 64, 64, 128),        # 4x4 patches of eight images per tile, 5 images: three boxes of the tile out of range
    (2, 24, 20, 128, 64, 64),        # ragged 24x20 map, unequal sources
    (1, 16, 16, 32, 64, 64),         # K chunks of 32 (SWIZZLE_64B)
    (16, 40, 40, 256, 256, 256),     # many rounds over the SMs
    (2, 160, 160, 64, 64, 64),       # wide patches (32 x 4)
]


@pytest.mark.parametrize("case", UP_CASES)
def test_conv_upsampled_source_equals_materialised_concat(case):
    """tcgen05 engine: conv(xu=low, x=skip) == conv(cat([upsample2x(low), skip])) bit for bit (same K order, same MMAs),
    and both agree with torch fp32 inside the bf16 conv tolerance."""
    Bn, H, W, Cu, Cx, Cout = case
    lib = L.lib()
    g = torch.Generator().manual_seed(sum(case))
    lo = torch.randn((Bn, H // 2, W // 2, Cu), generator=g).bfloat16()
    sk = torch.randn((Bn, H, W, Cx), generator=g).bfloat16()
    w = (torch.randn((Cout, 1, 1, Cu + Cx), generator=g) / (Cu + Cx) ** 0.5).bfloat16()
    bias = torch.randn((Cout,), generator=g)
    cat = torch.cat((F.interpolate(lo.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), sk.float().permute(0, 3, 1, 2)), 1)
    ref = F.silu(F.conv2d(cat, w.float().permute(0, 3, 1, 2), bias))
    lod, skd, wd, bd = lo.to(DEV), sk.to(DEV), w.to(DEV), bias.to(DEV)
    catd = cat.permute(0, 2, 3, 1).contiguous().bfloat16().to(DEV)
    y_mat = torch.full((Bn, H, W, Cout), 7.0, dtype=torch.bfloat16, device=DEV)
    y_up = torch.full((Bn, H, W, Cout), -7.0, dtype=torch.bfloat16, device=DEV)
    L.check(_conv(lib, catd, y_mat, wd, bd, 1, L.ENGINE_TCGEN05), "yre_conv")
    L.check(_conv(lib, skd, y_up, wd, bd, 1, L.ENGINE_TCGEN05, xu=lod), "yre_conv(xu)")
    torch.cuda.synchronize()
    assert torch.equal(y_mat, y_up), f"{case}: {(y_mat.float() - y_up.float()).abs().max().item()}"
    err = (y_up.float().cpu().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 1.0e-2 * max(1.0, ref.abs().max().item()), f"{case}: max err {err:.4f}"


def test_conv_upsampled_source_fp32_engine():
    """fp32 validation engine (SIMT FMA): same semantics, <= 1e-4 x max|ref| against torch fp32."""
    lib = L.lib()
    g = torch.Generator().manual_seed(5)
    Bn, H, W, Cu, Cx, Cout = 3, 12, 20, 48, 80, 72
    lo, sk = torch.randn((Bn, H // 2, W // 2, Cu), generator=g), torch.randn((Bn, H, W, Cx), generator=g)
    w = torch.randn((Cout, 1, 1, Cu + Cx), generator=g) / (Cu + Cx) ** 0.5
    bias = torch.randn((Cout,), generator=g)
    cat = torch.cat((F.interpolate(lo.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), sk.permute(0, 3, 1, 2)), 1)
    ref = F.silu(F.conv2d(cat, w.permute(0, 3, 1, 2), bias))
    y = torch.zeros((Bn, H, W, Cout), device=DEV)
    L.check(_conv(lib, sk.to(DEV), y, w.to(DEV), bias.to(DEV), 1, L.ENGINE_FFMA, xu=lo.to(DEV)), "yre_conv(xu, ffma)")
    err = (y.cpu().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), err


def test_conv_upsampled_source_rejections():
    """Shapes the engines do not take fail loudly: 3x3 kernels, a source that is not half the size, K chunks that would
    straddle the two sources on the tcgen05 engine."""
    lib = L.lib()
    z = lambda *s: torch.zeros(s, dtype=torch.bfloat16, device=DEV)
    b = torch.zeros((64,), device=DEV)
    assert _conv(lib, z(1, 16, 16, 64), z(1, 16, 16, 64), z(64, 1, 1, 128), b, 0, L.ENGINE_TCGEN05, xu=z(1, 8, 9, 64)) == -1
    assert _conv(lib, z(1, 16, 16, 96), z(1, 16, 16, 64), z(64, 1, 1, 128), b, 0, L.ENGINE_TCGEN05, xu=z(1, 8, 8, 32)) == -3
    x, y, xu = z(1, 16, 16, 64), z(1, 16, 16, 64), z(1, 8, 8, 64)
    d = L.ConvDesc(L.View(x.data_ptr(), L.BF16, L.NHWC, 1, 16, 16, 64, 0, 64), L.View(y.data_ptr(), L.BF16, L.NHWC, 1, 16, 16, 64, 0, 64),
                   _null(), z(64, 3, 3, 128).data_ptr(), b.data_ptr(), 3, 1, 0, L.ENGINE_AUTO, L.View(xu.data_ptr(), L.BF16, L.NHWC, 1, 8, 8, 64, 0, 64))
    assert lib.yre_conv(C.byref(d), None) == -3


@pytest.mark.parametrize("cfg,size,prec", [("gelan-c", 128, "bf16"), ("gelan-c", 320, "bf16"), ("gelan-c", 160, "fp32"), ("yolov9-c", 128, "bf16")])
def test_model_fused_upsample_equals_materialised(cfg, size, prec, request):
    """Whole model: the plan with the Upsample layers folded into their consumers (default) returns bit-identical outputs
    to the plan that runs K5 (upsample2x into the concat slice), with two launches fewer."""
    nodes, nc, sd = request.getfixturevalue("gelan_c" if cfg == "gelan-c" else "yolov9_c")
    x = G.fractal(3, size, torch.Generator().manual_seed(size)).to(DEV)
    outs, launches = [], []
    for fuse in (True, False):
        m = YOLO.from_yaml(ROOT / "configs/models" / f"{cfg}.yaml")
        m.load_state_dict(sd, strict=True)
        m = m.to(DEV).eval().set_precision(prec)
        m.fuse_upsample = fuse
        y, raws = m(x)
        torch.cuda.synchronize()
        plan = next(iter(m._plans.values()))
        names = [n for n, _ in plan.op_table()]
        launches.append(plan.num_launches)
        assert ("upsample2x" in names) == (not fuse)
        if prec == "bf16":
            assert "conv_ffma" not in names
        flat = lambda o: [o] if isinstance(o, torch.Tensor) else [t for q in o for t in flat(q)]
        outs.append([t.clone() for t in flat(y) + flat(raws)])
    assert launches[0] == launches[1] - 2
    for a, b in zip(*outs):
        assert torch.equal(a, b)


# ---- paired halo schedule for a single 64-wide N tile (head box tower: 3x3 256->64 on the 80x80 level) ----------------
@pytest.mark.parametrize("case", [(6, 80, 80, 256, 1, 0), (15, 48, 56, 128, 1, 1), (5, 80, 80, 192, 0, 0),
                                  (24, 40, 40, 128, 1, 1), (42, 24, 40, 192, 1, 0), (64, 40, 40, 512, 0, 0)])
def test_conv_halo_pair_cout64(case):
    """3x3 stride-1, Cin a multiple of 64, Cout = 64, enough exact 8x16 patches: two patches share every weight box
    (TcParams::npair == 2 with a 64-column N tile); even / odd patch counts, residual, no activation.
    The 40-wide maps are not tiled by 8x16 patches: they take 8x8 patches of two images (TcParams::ybx, tensor maps with
    the batch dimension ahead of H) -- even / odd pair counts (300, 315 tiles), residual, the head's 3x3 512->64 @40x40."""
    Bn, H, W, Cin, act, res = case
    lib = L.lib()
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn((Bn, H, W, Cin), generator=g).bfloat16()
    w = (torch.randn((64, 3, 3, Cin), generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn((64,), generator=g)
    r = torch.randn((Bn, H, W, 64), generator=g).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    if act:
        ref = F.silu(ref)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    xd, wd, bd, rd = x.to(DEV), w.to(DEV), bias.to(DEV), r.to(DEV)
    y = torch.full((Bn, H, W, 64), 7.0, dtype=torch.bfloat16, device=DEV)
    d = L.ConvDesc(L.View(xd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cin, 0, Cin), L.View(y.data_ptr(), L.BF16, L.NHWC, Bn, H, W, 64, 0, 64),
                   L.View(rd.data_ptr(), L.BF16, L.NHWC, Bn, H, W, 64, 0, 64) if res else _null(),
                   wd.data_ptr(), bd.data_ptr(), 3, 1, act, L.ENGINE_TCGEN05)
    L.check(lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream), "yre_conv")
    err = (y.float().cpu().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 1.0e-2 * max(1.0, ref.abs().max().item()), f"{case}: max err {err:.4f}"
