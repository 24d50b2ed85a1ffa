"""Runs the step's top conv shapes (batch 64) once each through the C ABI, for an `ncu --set full --import-source on` capture:
   ncu --set full --import-source on --clock-control none -k regex:conv -o gpurun_out/r2_prof_top python scripts/ncu_top_convs.py
Order: 3x3 256->256 @80x80 (CTA-pair generic), 1x1 1024->512 @40x40 (CTA-pair generic), 3x3 128->128 @80x80 (paired halo stream),
3x3 64->64 @80x80 (weight-stationary halo), 1x1 256->256 @160x160 (HBM-bound, CTA pairs + 64-column staging),
3x3 128->128 @40x40 (CTA-pair generic)."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from yolo_b200 import _lib as L

SHAPES = [(64, 80, 80, 256, 256, 3), (64, 40, 40, 1024, 512, 1), (64, 80, 80, 128, 128, 3), (64, 80, 80, 64, 64, 3),
          (64, 160, 160, 256, 256, 1), (64, 40, 40, 128, 128, 3)]
lib = L.lib()
s = torch.cuda.current_stream().cuda_stream
for Bn, H, W, Cin, Cout, k in SHAPES:
    x = torch.randn((Bn, H, W, Cin), device="cuda").bfloat16()
    w = (torch.randn((Cout, k, k, Cin), device="cuda") / (k * k * Cin) ** 0.5).bfloat16()
    bias = torch.zeros((Cout,), device="cuda")
    y = torch.empty((Bn, H, W, Cout), device="cuda", dtype=torch.bfloat16)
    d = L.ConvDesc(L.View(x.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cin, 0, Cin), L.View(y.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cout, 0, Cout),
                   L.View(None, 0, 0, 0, 0, 0, 0, 0, 0), w.data_ptr(), bias.data_ptr(), k, 1, 1, L.ENGINE_TCGEN05)
    L.check(lib.yre_conv(C.byref(d), s), "yre_conv")
    torch.cuda.synchronize()
print("done")
