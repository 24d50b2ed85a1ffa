#!/bin/bash
# Round-2 session 3: virtual Upsample+Concat input (yre_conv_desc.xu) -- tests first, then the per-op bench.
mkdir -p gpurun_out
echo "== upcat tests"; timeout 900 python -m pytest tests/test_gpu_upcat.py -x -q -p no:cacheprovider > gpurun_out/t3_up.log 2>&1; tail -25 gpurun_out/t3_up.log | cut -c1-300
echo "== bench config 2"; timeout 900 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --per-op gpurun_out/per_op3_c2.csv > gpurun_out/bench3_c2.log 2>gpurun_out/bench3_c2.err; tail -1 gpurun_out/bench3_c2.log | cut -c1-400; tail -3 gpurun_out/bench3_c2.err
