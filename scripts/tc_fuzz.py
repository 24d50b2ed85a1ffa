"""GPU-box fuzz: random conv shapes, tcgen05 engines (generic / halo / halo-stream / pairs) against the fp32 FFMA
engine on bf16-rounded operands.  Every case runs in this process; a sticky CUDA error aborts the run.

    python scripts/tc_fuzz.py [n_cases] [seed] [pairs]
"""
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "scripts"))
sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import tc_diag  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for it in range(n):
    k = rnd.choice([1, 3, 3])
    stride = 2 if (k == 3 and rnd.random() < 0.2) else 1
    Cin = rnd.choice([32, 64, 64, 128, 256])
    Cout = rnd.choice([32, 64, 64, 128, 128, 160, 256, 320])
    extra = rnd.choice([0, 0, 32, 64])
    Ct = Cin + extra
    coff = rnd.choice([0, extra]) if extra else 0
    if rnd.random() < 0.5:      # sizes that tile into 8x16 patches -> halo kernels (and pairs when Cout % 128 == 0)
        H, W = 16 * rnd.randint(1, 5), 8 * rnd.randint(1, 10)
    else:
        H, W = rnd.randint(3, 70), rnd.randint(3, 70)
    Bn = rnd.choice([1, 2, 3, 5, 8])
    if Bn * H * W * max(Ct, Cout) > 40_000_000:
        Bn = 1
    if len(sys.argv) > 3 and sys.argv[3] == "pairs":     # enough 8x16 patches for the paired halo-stream schedule (>= 148 units)
        k, stride, Cin, Cout = 3, 1, rnd.choice([64, 128, 256]), rnd.choice([128, 256])
        Ct, coff = Cin + extra, (rnd.choice([0, extra]) if extra else 0)
        H, W = 16 * rnd.randint(2, 5), 8 * rnd.randint(3, 10)
        Bn = max(1, -(-300 // ((H // 16) * (W // 8)))) + rnd.randint(0, 3)
    case = (Bn, H, W, Ct, coff, Cin, Cout, k, stride, rnd.choice([0, 1]), rnd.choice([0, 1]), 0)
    tc_diag.ALL.append(case)
    rc = tc_diag.run_case(len(tc_diag.ALL) - 1)
    if rc:
        bad += 1
        if rc == 1:
            break
print(f"fuzz: {n} cases, {bad} failures")
sys.exit(1 if bad else 0)
