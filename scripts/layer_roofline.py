"""Per-layer roofline table from a bench --per-op CSV (offline, no GPU needed).

    python scripts/layer_roofline.py profiles/r01_per_op.csv profiles/r01_layer_roofline.csv

For every conv launch: algorithmic FLOPs (from the CSV), algorithmic bytes 2*(|in| + |out| + |w|) in bf16 (fp32 outputs
counted at 4 B), the ideal time max(FLOPs / sustained bf16 peak, bytes / measured HBM bandwidth) and the fraction of that
ideal the measured launch reaches.  Peaks: MEASURED_PEAKS.json if present, else 1374.1 TFLOP/s and 6549.8 GB/s."""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
src, dst = Path(sys.argv[1]), Path(sys.argv[2])
tf, gbs = 1374.1, 6549.8
mp = ROOT / "MEASURED_PEAKS.json"
if mp.exists():
    d = json.loads(mp.read_text())
    tf, gbs = d.get("bf16_tflops_sustained", tf), d.get("hbm_gbs", gbs)
rows = list(csv.DictReader(open(src)))
out = []
pat = re.compile(r"conv(\d)x\d+s(\d) (\d+)->(\d+) @(\d+)x(\d+) B(\d+)( \+res)?( f32out)?")
for r in rows:
    m = pat.match(r["shape"])
    if r["kernel"] != "conv_tc" or not m:
        continue
    k, s, cin, cout, h, w, b = (int(m.group(i)) for i in range(1, 8))
    res, f32 = bool(m.group(8)), bool(m.group(9))
    px_out = b * h * w
    px_in = px_out * s * s
    bytes_ = 2 * px_in * cin + (4 if f32 else 2) * px_out * cout + 2 * cout * cin * k * k + (2 * px_out * cout if res else 0)
    flops = float(r["gflop"]) * 1e9
    t = float(r["ms"]) * 1e-3
    t_tensor, t_hbm = flops / (tf * 1e12), bytes_ / (gbs * 1e9)
    ideal = max(t_tensor, t_hbm)
    out.append((r["op"], r["shape"], f"{t * 1e6:.1f}", f"{flops / t / 1e12:.0f}", f"{bytes_ / 1e6:.1f}", f"{bytes_ / t / 1e9:.0f}",
                "tensor" if t_tensor >= t_hbm else "hbm", f"{ideal * 1e6:.1f}", f"{ideal / t:.2f}"))
with open(dst, "w") as f:
    f.write("op,shape,time_us,tflops,algorithmic_MB,GBps,bound,ideal_us,frac_of_ideal\n")
    for o in out:
        f.write(",".join(o) + "\n")
tot_t = sum(float(o[2]) for o in out)
tot_i = sum(float(o[7]) for o in out)
print(f"{len(out)} conv launches: measured {tot_t / 1e3:.3f} ms, sum of per-layer ideals {tot_i / 1e3:.3f} ms -> {tot_i / tot_t:.2f} of the unfused per-layer roofline")
