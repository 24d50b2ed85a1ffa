"""Diagnostic (GPU box only): runs one conv shape through the tcgen05 engine and the FFMA engine via
the C ABI and prints how they differ.  Each case runs in its own process because a trapped kernel
poisons the CUDA context.   python scripts/tc_diag.py [case index | all]"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "yolo-re_b200"))

# (B, H, W, Cin_total, c_off, Cin, Cout, k, stride(phase4), act, res, out_f32)
CASES = [
    (1, 8, 16, 64, 0, 64, 64, 1, 1, 0, 0, 0),       # one tile, one k-iter, no epilogue math
    (1, 8, 16, 64, 0, 64, 64, 1, 1, 1, 0, 0),       # + SiLU
    (1, 8, 16, 128, 0, 128, 128, 1, 1, 0, 0, 0),    # 2 k-iters
    (1, 8, 16, 64, 0, 64, 64, 3, 1, 0, 0, 0),       # 3x3 halo via OOB fill
    (2, 20, 20, 256, 128, 128, 128, 3, 1, 1, 1, 0),  # channel window, residual, ragged tiles, batch tiling
    (1, 24, 24, 32, 0, 32, 32, 3, 1, 1, 0, 0),      # BLOCK_K=32 / SW64
    (1, 16, 16, 256, 0, 256, 80, 1, 1, 0, 0, 1),    # N=80, fp32 out
    (2, 40, 40, 512, 0, 512, 512, 1, 1, 1, 0, 0),   # 2 N tiles, many tiles (persistent loop, TMEM double buffer)
    (2, 31, 31, 64, 0, 64, 64, 3, 2, 1, 0, 0),      # stride 2 from PHASE4 (odd extent)
    (1, 80, 80, 256, 0, 256, 320, 3, 1, 1, 0, 0),   # N=320 -> 2x160
]


BIG = [
    (64, 80, 80, 256, 0, 256, 320, 3, 1, 1, 0, 0),    # bn=160 -> 3 accumulators
    (64, 40, 40, 256, 0, 256, 256, 3, 1, 1, 0, 0),    # bn=256 -> 2 accumulators, single pipe
    (64, 160, 160, 64, 0, 64, 64, 3, 1, 1, 1, 0),     # many tiles, residual, dual pipe
    (64, 159, 159, 128, 0, 128, 128, 3, 2, 1, 0, 0),  # stride 2 from parity planes
    (64, 160, 160, 32, 0, 32, 32, 3, 1, 1, 0, 0),     # BLOCK_K=32
    (64, 40, 40, 1024, 0, 1024, 512, 1, 1, 1, 0, 0),  # 2 n-tiles of 256
    (64, 80, 80, 256, 0, 256, 80, 1, 1, 0, 0, 1),     # fp32 direct-store head conv
    (64, 160, 160, 64, 0, 32, 32, 3, 1, 1, 0, 0),     # channel window of a wider buffer, BLOCK_K=32
    (16, 160, 160, 64, 0, 32, 32, 3, 1, 1, 0, 0),
    (64, 160, 160, 64, 32, 32, 32, 3, 1, 1, 0, 0),
    (5, 80, 72, 128, 0, 128, 256, 3, 1, 1, 1, 0),     # 20: halo-stream pairs, odd patch count (last pair half empty), residual
    (64, 80, 80, 128, 0, 128, 128, 3, 1, 1, 0, 0),    # 21: halo-stream pairs
    (64, 80, 80, 256, 0, 256, 256, 3, 1, 1, 1, 0),    # 22: halo-stream pairs, 2 N tiles, 4 channel chunks
]


def run_case(i):
    import torch
    from yolo_b200 import _lib as L
    lib = L.lib()
    Bn, H, W, Ct, coff, Cin, Cout, k, stride, act, res, of32 = ALL[i]
    g = torch.Generator().manual_seed(i)
    dev = "cuda"
    if stride == 2:
        Hp, Wp = (H + 1) // 2, (W + 1) // 2
        x = torch.zeros((4, Bn, Hp, Wp, Ct), dtype=torch.bfloat16)
        full = torch.randn((Bn, H, W, Ct), generator=g).bfloat16()
        for py in range(2):
            for px in range(2):
                sub = full[:, py::2, px::2]
                x[py * 2 + px, :, :sub.shape[1], :sub.shape[2]] = sub
        layout = L.PHASE4
        Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    else:
        x = torch.randn((Bn, H, W, Ct), generator=g).bfloat16()
        layout = L.NHWC
        Ho, Wo = H, W
    x = x.to(dev)
    w = (torch.randn((Cout, k, k, Cin), generator=g) / (k * k * Cin) ** 0.5).bfloat16().to(dev)
    bias = torch.randn((Cout,), generator=g).to(dev)
    r = torch.randn((Bn, Ho, Wo, Cout), generator=g).bfloat16().to(dev)
    outs = []
    for eng in (L.ENGINE_FFMA, L.ENGINE_TCGEN05):
        y = torch.full((Bn, Ho, Wo, Cout), 7.0, dtype=torch.float32 if of32 else torch.bfloat16, device=dev)
        d = L.ConvDesc(L.View(x.data_ptr(), L.BF16, layout, Bn, H, W, Ct, coff, Cin),
                       L.View(y.data_ptr(), L.F32 if of32 else L.BF16, L.NHWC, Bn, Ho, Wo, Cout, 0, Cout),
                       L.View(r.data_ptr(), L.BF16, L.NHWC, Bn, Ho, Wo, Cout, 0, Cout) if res else L.View(None, 0, 0, 0, 0, 0, 0, 0, 0),
                       w.data_ptr(), bias.data_ptr(), k, stride, act, eng)
        rc = lib.yre_conv(C.byref(d), torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            print(f"case {i}: engine {eng} rc={rc}: {lib.yre_last_error().decode()}")
            return 1
        torch.cuda.synchronize()
        outs.append(y.float().cpu())
    a, b = outs
    err = (a - b).abs()
    tol = 2e-2 * max(1.0, a.abs().max().item())
    bad = err > tol
    print(f"case {i} {ALL[i]}: max|ffma|={a.abs().max():.3f} max err={err.max():.4f} bad={bad.float().mean():.4f} "
          f"untouched(7.0)={(b == 7.0).float().mean():.4f}")
    if bad.any():
        idx = bad.nonzero()
        print("   first bad (b,y,x,c):", idx[:6].tolist())
        print("   bad frac by channel%8:", [round(bad[..., c::8].float().mean().item(), 3) for c in range(8)])
        print("   bad frac by x%8:", [round(bad[:, :, xx::8].float().mean().item(), 3) for xx in range(min(8, Wo))])
        print("   bad frac by 16-ch group:", [round(bad[..., c:c + 16].float().mean().item(), 3) for c in range(0, min(Cout, 128), 16)])
        print("   sample ffma:", a[tuple(idx[0])].item(), "tc:", b[tuple(idx[0])].item())
        return 2
    return 0


ALL = CASES + BIG

if __name__ == "__main__":
    arg = sys.argv[1] if len(sys.argv) > 1 else "all"
    if arg in ("all", "big"):
        rcs = []
        for i in (range(len(CASES)) if arg == "all" else range(len(CASES), len(ALL))):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], timeout=120, capture_output=True, text=True)
                print(r.stdout.strip() or f"case {i}: no output", flush=True)
                if r.returncode not in (0, 2):
                    print(f"   case {i} exit {r.returncode}: {r.stderr.strip()[-600:]}", flush=True)
                rcs.append(r.returncode)
            except subprocess.TimeoutExpired:
                print(f"case {i}: TIMEOUT", flush=True)
                rcs.append(-9)
        print("summary:", rcs)
    else:
        sys.exit(run_case(int(arg)))
