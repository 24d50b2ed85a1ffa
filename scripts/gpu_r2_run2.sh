#!/bin/bash
# Round-2 GPU session 2: new NMS select kernel + selective CTA pairs; tests first, then benches and tuning experiments.
mkdir -p gpurun_out
echo "== nms + round2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_ops.py -q -p no:cacheprovider -s > gpurun_out/t2_a.log 2>&1; tail -30 gpurun_out/t2_a.log | cut -c1-260
echo "== model tests"; timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_preproc.py tests/test_gpu_metrics.py -q -p no:cacheprovider > gpurun_out/t2_b.log 2>&1; tail -5 gpurun_out/t2_b.log | cut -c1-260
echo "== bench config 2"; timeout 900 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --per-op gpurun_out/per_op2_c2.csv > gpurun_out/bench2_c2.log 2>gpurun_out/bench2_c2.err; tail -1 gpurun_out/bench2_c2.log | cut -c1-200; tail -3 gpurun_out/bench2_c2.err
echo "== bench config 5"; timeout 900 python bench.py --config 5 --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench2_c5.log 2>gpurun_out/bench2_c5.err; tail -1 gpurun_out/bench2_c5.log | cut -c1-200; tail -3 gpurun_out/bench2_c5.err
for v in "YRE_TC_BLOCK_N=128" "YRE_TC_CTA2=0" "YRE_TC_STAGES=3"; do
  echo "== per-op with $v"; env $v timeout 600 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline --per-op "gpurun_out/per_op2_${v}.csv" > "gpurun_out/bench2_${v}.log" 2>&1; tail -1 "gpurun_out/bench2_${v}.log" | cut -c1-160
done
echo "== traces"; timeout 600 python scripts/tc_trace.py mem > gpurun_out/trace_mem.log 2>&1; tail -40 gpurun_out/trace_mem.log | cut -c1-2500
