"""GPU-box diagnostic: per-role timeline of CTA 0 of the tcgen05 conv (YRE_TC_TRACE=1)."""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["YRE_TC_TRACE"] = "1"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from yolo_b200 import _lib as L

lib = L.lib()
lib.yre_debug_read_trace.argtypes = [C.c_void_p, C.c_int]
lib.yre_debug_read_trace.restype = C.c_int
NAMES = {1: "P:empty-ok", 2: "P:tma-issued", 10: "M:tmem-free", 11: "M:full-ok", 12: "M:mma-issued", 20: "E:start", 21: "E:acc-ready", 22: "E:done", 23: "e:ld", 24: "e:math", 25: "e:bufwait", 26: "e:stored"}


def run(Bn, H, W, Cin, Cout, k):
    dev = "cuda"
    x = torch.randn((Bn, H, W, Cin), device=dev).bfloat16()
    w = torch.randn((Cout, k, k, Cin), device=dev).bfloat16()
    bias = torch.zeros((Cout,), device=dev)
    y = torch.empty((Bn, H, W, Cout), device=dev, dtype=torch.bfloat16)
    d = L.ConvDesc(L.View(x.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cin, 0, Cin), L.View(y.data_ptr(), L.BF16, L.NHWC, Bn, H, W, Cout, 0, Cout),
                   L.View(None, 0, 0, 0, 0, 0, 0, 0, 0), w.data_ptr(), bias.data_ptr(), k, 1, 1, L.ENGINE_TCGEN05)
    for _ in range(3):
        L.check(lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_int * 8192)()
    lib.yre_debug_read_trace(buf, 8192)
    print(f"\n=== conv{k}x{k} {Cin}->{Cout} @{H}x{W} B{Bn}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    import struct
    ent = []
    for b in range(148):
        o = 2048 + b * 8
        g0 = ((buf[o + 1] & 0xffffffff) << 32) | (buf[o] & 0xffffffff); c0 = ((buf[o + 3] & 0xffffffff) << 32) | (buf[o + 2] & 0xffffffff)
        g1 = ((buf[o + 5] & 0xffffffff) << 32) | (buf[o + 4] & 0xffffffff); c1 = ((buf[o + 7] & 0xffffffff) << 32) | (buf[o + 6] & 0xffffffff)
        if g0 and g1: ent.append((g0, g1, c1 - c0))
    if ent:
        gmin = min(e[0] for e in ent); gmax = max(e[1] for e in ent)
        durs = sorted(e[1] - e[0] for e in ent)
        print(f" CTAs={len(ent)} grid span={gmax - gmin} ns; entry spread={max(e[0] for e in ent) - gmin} ns; CTA dur min/med/max={durs[0]}/{durs[len(durs)//2]}/{durs[-1]} ns;"
              f" clk/ns med={sorted(e[2] / max(1, e[1] - e[0]) for e in ent)[len(ent)//2]:.3f}")
    evs = []
    for role in range(3):
        for n in range(96):
            s = 16 + (role * 96 + n) * 4
            if buf[s + 3]:
                evs.append((((buf[s + 2] & 0xffffffff) << 32) | (buf[s + 1] & 0xffffffff), role, buf[s]))
    if not evs:
        print("no trace")
        return
    t0 = min(e[0] for e in evs)
    for role in range(3):
        seq = [(t - t0, ev) for t, r, ev in evs if r == role]
        seq.sort()
        line = []
        prev = seq[-61][0] if len(seq) > 60 else 0
        for t, ev in seq[-60:]:
            line.append(f"{NAMES[ev]}@{t}(+{t - prev})")
            prev = t
        print(" role", role, " ".join(line))


shapes = [(8, 160, 160, 32, 32, 3), (8, 160, 160, 64, 64, 1), (8, 80, 80, 128, 128, 3), (8, 80, 80, 256, 256, 3)]
if len(sys.argv) > 1 and sys.argv[1] == 'small':
    shapes = [(64, 20, 20, 256, 256, 1), (64, 20, 20, 128, 128, 3), (64, 20, 20, 64, 64, 3)]
elif len(sys.argv) > 1 and sys.argv[1] == 'halo':
    shapes = [(64, 80, 80, 64, 64, 3), (64, 160, 160, 32, 32, 3), (64, 160, 160, 64, 64, 3)]
elif len(sys.argv) > 1 and sys.argv[1] == 'mem':
    shapes = [(64, 80, 80, 128, 128, 1), (64, 40, 40, 256, 256, 1), (64, 40, 40, 1024, 512, 1)]
for shape in shapes:
    run(*shape)
