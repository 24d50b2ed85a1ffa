#!/bin/bash
# ncu: launch list of our kernels for one step + full captures of the bandwidth kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc|conv_ffma|stem|adown|spp|upsample|decode|nms|cbfuse" -s 411 -c 137 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"stem|adown|decode|upsample" -s 27 -c 9 -o gpurun_out/prof_bw \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bw.log 2>&1
ls -la gpurun_out
