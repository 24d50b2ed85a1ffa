#!/bin/bash
# End-of-round GPU session: parity suite, smoke, bench (+ per-op table), then the ncu evidence for profiles/:
# launch list with DRAM bytes, speed-of-light section of every conv launch of one step, one full capture with source.
# Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
K='regex:conv|stem|adown|spp|upsample|decode|nms|cbfuse'
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --per-op gpurun_out/per_op.csv > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-300
cp gpurun_out/bench.log gpurun_out/bench_n1.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
L=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/plain.log') if l.startswith('{')][-1])['launches_per_step'])")
echo "== ncu launch list ($L launches per step)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s $((3 * L)) -c $L --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
NC=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/plain.log') if l.startswith('{')][-1])['tcgen05_convs_per_step'])")
echo "== ncu SpeedOfLight of the $NC conv launches of one step"
timeout 1500 ncu --section SpeedOfLight --clock-control none -k regex:conv -s $((3 * NC)) -c $NC -o gpurun_out/prof_conv_all -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "== ncu full + source: paired halo-stream conv, generic conv, weight-stationary halo conv"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv3_halo_stream -s 12 -c 1 -o gpurun_out/prof_src_stream -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 150 -c 1 -o gpurun_out/prof_src_generic -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:conv3_halo_kernel" -s 30 -c 1 -o gpurun_out/prof_src_halo -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -20
