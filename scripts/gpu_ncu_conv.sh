#!/bin/bash
# ncu section capture of every conv_tc launch of one bench step (run only after the same command exited 0 without ncu).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section LaunchStats --section Occupancy \
    --clock-control none -k regex:conv_tc -s 372 -c 124 -o gpurun_out/prof_conv_all -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log; ls -la gpurun_out/prof_conv_all.ncu-rep
