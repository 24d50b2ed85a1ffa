#!/bin/bash
# GPU-box session: full parity suite, bench with per-op table, ncu launch list + conv_tc captures.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --per-op gpurun_out/per_op.csv > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -2 gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  echo "== ncu launch list"
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_tc|conv_ffma|stem|adown|spp|upsample|decode|nms|cbfuse" -s 411 -c 274 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
  echo "== ncu conv_tc sections (one step, all 124 launches)"
  timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy \
      --clock-control none -k regex:conv_tc -s 372 -c 124 -o gpurun_out/prof_conv_all \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
  echo "== ncu conv_tc full + source (3 launches)"
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 480 -c 3 -o gpurun_out/prof_conv_src \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
  ls -la gpurun_out
fi
