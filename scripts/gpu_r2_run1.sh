#!/bin/bash
# Round-2 GPU session 1: new-feature parity first (CTA pairs isolated in their own process), then the suite, smoke,
# and the bench at configs 2 / 4 / 5 with per-op tables; CTA pairs on vs off (tuning build) for the A/B.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader | tee gpurun_out/gpu.txt
echo "== pair tests"; timeout 600 python -m pytest tests/test_gpu_round2.py -k "cta_pair" -q -p no:cacheprovider > gpurun_out/t_pair.log 2>&1; rc=$?
tail -15 gpurun_out/t_pair.log
if [ $rc -ne 0 ]; then echo "PAIR TESTS FAILED (rc=$rc): continuing with YRE_TC_CTA2=0"; export YRE_TC_CTA2=0; fi
echo "== round2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py -k "not cta_pair" -q -p no:cacheprovider -s > gpurun_out/t_round2.log 2>&1; tail -25 gpurun_out/t_round2.log | cut -c1-300
echo "== ops tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider > gpurun_out/t_ops.log 2>&1; tail -8 gpurun_out/t_ops.log | cut -c1-300
echo "== model tests"; timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_preproc.py tests/test_gpu_metrics.py -q -p no:cacheprovider -s > gpurun_out/t_model.log 2>&1; tail -12 gpurun_out/t_model.log | cut -c1-300
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench config 2"; timeout 900 python bench.py --steps 20 --warmup 5 --per-op gpurun_out/per_op_c2.csv > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; tail -1 gpurun_out/bench_c2.log | cut -c1-400; tail -3 gpurun_out/bench_c2.err
echo "== bench config 2, CTA pairs off"; YRE_TC_CTA2=0 timeout 900 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --per-op gpurun_out/per_op_c2_nopair.csv > gpurun_out/bench_c2_nopair.log 2>gpurun_out/bench_c2_nopair.err; tail -1 gpurun_out/bench_c2_nopair.log | cut -c1-300
echo "== bench config 5"; timeout 900 python bench.py --config 5 --steps 10 --warmup 3 --per-op gpurun_out/per_op_c5.csv > gpurun_out/bench_c5.log 2>gpurun_out/bench_c5.err; tail -1 gpurun_out/bench_c5.log | cut -c1-400; tail -3 gpurun_out/bench_c5.err
echo "== bench config 4"; timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --per-op gpurun_out/per_op_c4.csv > gpurun_out/bench_c4.log 2>gpurun_out/bench_c4.err; tail -1 gpurun_out/bench_c4.log | cut -c1-400; tail -3 gpurun_out/bench_c4.err
echo "== bench config 4 main-only"; timeout 900 python bench.py --config 4 --main-only --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_c4_main.log 2>gpurun_out/bench_c4_main.err; tail -1 gpurun_out/bench_c4_main.log | cut -c1-300
ls gpurun_out | wc -l
