"""Turns the raw artefacts a GPU session left in gpurun_out/ into the small, tracked summaries under
profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch).

    python scripts/summarize_profiles.py r01            # round-1 file names (bench.log, launches.csv, ...)
    python scripts/summarize_profiles.py r02 r2_        # round-2 session (scripts/gpu_r2_final.sh): gpurun_out/r2_*
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
pre = sys.argv[2] if len(sys.argv) > 2 else ""
P.mkdir(exist_ok=True)

# 1. bench line + per-op table
for name in ("bench.log", "bench_n1.log", "bench_n2.log", "bench_n4.log", "bench_n8.log", "bench_c3_n1.log", "bench_c3_n2.log", "bench_c3_n4.log",
             "bench_c3_n8.log", "bench_c4.log", "bench_c4_main.log", "bench_c4_n8.log", "bench_c5.log", "bench_c5_n8.log", "bench_reference.log"):
    f = G / (pre + name)
    if f.exists():
        lines = [l for l in f.read_text().splitlines() if l.startswith("{")]
        if lines:
            (P / f"{tag}_{name.replace('.log', '.json')}").write_text(json.dumps(json.loads(lines[-1]), indent=1) + "\n")
for name in ("per_op.csv", "per_op_c4.csv", "per_op_c5.csv"):
    if (G / (pre + name)).exists():
        shutil.copy(G / (pre + name), P / f"{tag}_{name}")

# 2. ncu launch list (gpu__time_duration + dram bytes per launch) -> per-kernel summary
f = G / (pre + "launches.csv")
if f.exists():
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, iv, im, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > iv:
            d.setdefault(r[iid], {"k": r[ik].split("(")[0].split("::")[-1].replace(", ", ";")})[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.OrderedDict()
    for v in d.values():
        a = agg.setdefault(v["k"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += v.get("gpu__time_duration.sum", 0); a[2] += v.get("dram__bytes_read.sum", 0); a[3] += v.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    with open(P / f"{tag}_ncu_launch_summary.csv", "w") as o:
        o.write("kernel,launches_per_step,time_us,share_of_step,dram_read_MB,dram_write_MB,dram_GBps\n")
        for k, (n, t, r, w) in agg.items():
            o.write(f"{k},{n},{t / 1e3:.1f},{t / tot:.4f},{r / 1e6:.1f},{w / 1e6:.1f},{(r + w) / t if t else 0:.1f}\n")
    with open(P / f"{tag}_ncu_launches.csv", "w") as o:
        o.write("id,kernel,time_us,dram_read_MB,dram_write_MB\n")
        for i, v in d.items():
            o.write(f"{i},{v['k']},{v.get('gpu__time_duration.sum', 0) / 1e3:.2f},{v.get('dram__bytes_read.sum', 0) / 1e6:.2f},{v.get('dram__bytes_write.sum', 0) / 1e6:.2f}\n")

# 3. speed-of-light table of every conv_tc launch of one step (from an ncu --section capture)
rep = G / (pre + "prof_conv_all.ncu-rep")
if rep.exists():
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[0], rows[2:]
    cols = {"time_us": "gpu__time_duration.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed"}
    idx = {k: hdr.index(v) for k, v in cols.items() if v in hdr}
    shapes = [r["shape"] for r in csv.DictReader(open(G / (pre + "per_op.csv"))) if r["kernel"] == "conv_tc"] if (G / (pre + "per_op.csv")).exists() else []
    with open(P / f"{tag}_ncu_conv_tc_sol.csv", "w") as o:
        o.write("launch,shape," + ",".join(idx) + "\n")
        for i, r in enumerate(data):
            o.write(f"{i},{shapes[i] if i < len(shapes) else ''}," + ",".join(r[j].replace(",", "") for j in idx.values()) + "\n")
print("profiles/:", sorted(p.name for p in P.iterdir()))
