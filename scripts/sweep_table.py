"""Merges the per-op CSVs of scripts/gpu_sweep.sh: per layer the time of every variant and the best one.
python scripts/sweep_table.py [dir] > table"""
import csv, glob, os, sys
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
runs = {}
for f in sorted(glob.glob(os.path.join(d, "sweep_*.csv"))):
    name = os.path.basename(f)[6:-4]
    runs[name] = list(csv.DictReader(open(f)))
names = list(runs)
base = runs["base"]
tot = {n: 0.0 for n in names}
best_tot = 0.0
print("op,shape,base_us," + ",".join(n for n in names if n != "base") + ",best,gain_us")
for i, r in enumerate(base):
    if not r["kernel"].startswith("conv"):
        continue
    row = {}
    for n in names:
        rr = runs[n]
        if i < len(rr) and rr[i]["shape"] == r["shape"]:
            row[n] = float(rr[i]["ms"]) * 1e3
    b0 = min(row.get("base", 1e9), row.get("base2", 1e9))
    bn = min(row, key=row.get)
    for n in names:
        tot[n] += row.get(n, b0)
    best_tot += row[bn]
    print(f"{r['op']},{r['shape']},{b0:.1f}," + ",".join(f"{row.get(n, float('nan')):.1f}" for n in names if n != "base") + f",{bn},{b0 - row[bn]:.1f}")
print("# totals (us):", {n: round(v) for n, v in tot.items()}, "best-per-layer:", round(best_tot))
