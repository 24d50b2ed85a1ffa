#!/bin/bash
# Round-2 evidence session (product build): parity suite, smoke, the bench at BASELINE configs 2 / 4 / 5 (+ per-op tables),
# then the ncu evidence for profiles/: launch list with DRAM bytes, SpeedOfLight of every conv launch of one step, one full
# capture of the CTA-pair conv kernel.  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
K='regex:conv|stem|adown|spp|upsample|decode|nms|cbfuse'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader | tee gpurun_out/gpu.txt
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
echo "== bench config 2"; timeout 900 python bench.py --steps 20 --warmup 5 --per-op gpurun_out/r2_per_op.csv > gpurun_out/r2_bench.log 2>gpurun_out/r2_bench.err; tail -1 gpurun_out/r2_bench.log | cut -c1-300; tail -2 gpurun_out/r2_bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.log 2>&1; tail -1 gpurun_out/r2_bench_reference.log | cut -c1-200
echo "== bench config 5"; timeout 900 python bench.py --config 5 --steps 10 --warmup 3 --per-op gpurun_out/r2_per_op_c5.csv > gpurun_out/r2_bench_c5.log 2>gpurun_out/r2_bench_c5.err; tail -1 gpurun_out/r2_bench_c5.log | cut -c1-200
echo "== bench config 4"; timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --per-op gpurun_out/r2_per_op_c4.csv > gpurun_out/r2_bench_c4.log 2>gpurun_out/r2_bench_c4.err; tail -1 gpurun_out/r2_bench_c4.log | cut -c1-200
echo "== bench config 4 main-only"; timeout 900 python bench.py --config 4 --main-only --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_main.log 2>&1; tail -1 gpurun_out/r2_bench_c4_main.log | cut -c1-200
echo "== bench config 3 at N=1 (512 images on one GPU)"; timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2_bench_c3_n1.log 2>&1; tail -1 gpurun_out/r2_bench_c3_n1.log | cut -c1-200
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_plain.log 2>&1 || { tail -5 gpurun_out/r2_plain.log; exit 1; }
L=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/r2_plain.log') if l.startswith('{')][-1])['launches_per_step'])")
echo "== ncu launch list ($L launches per step)"
# the quick bench runs: 3 warm-up + 2 (resident warm-up) + 2 timed public + 2 no-sync steps before the e2e legs; take one step of the timed loop
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s $((5 * L)) -c $L --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu1.log 2>&1
NC=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/r2_plain.log') if l.startswith('{')][-1])['tcgen05_convs_per_step'])")
echo "== ncu SpeedOfLight of the $NC conv launches of one step"
timeout 1500 ncu --section SpeedOfLight --clock-control none -k regex:conv -s $((5 * NC)) -c $NC -o gpurun_out/r2_prof_conv_all -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu2.log 2>&1
echo "== ncu full + source: CTA-pair generic conv (1x1 1024->512 @40x40), nms_select"
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:conv_tc_kernel.*true|conv_tc_kernel<.*1>" -s 40 -c 1 -o gpurun_out/r2_prof_src_pair -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:nms_select -s 6 -c 1 -o gpurun_out/r2_prof_src_nms -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r2_ncu4.log 2>&1
echo "== sanitizers"; bash scripts/gpu_sanitize.sh 2>&1 | tail -12
ls -la gpurun_out | grep r2_ | wc -l
