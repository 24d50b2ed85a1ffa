#!/bin/bash
# Per-layer sweep at another BASELINE configuration (fewer variants).  usage: gpu_sweep3.sh TUNE.so CONFIG
mkdir -p gpurun_out
VAR=$1; CFG=${2:-4}
LIB=yolo-re_b200/yolo_b200/libyre.so
cp $LIB /tmp/base.so; cp $VAR $LIB
rm -f gpurun_out/sweep_*.csv gpurun_out/sweep_*.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --config $CFG --steps 5 --warmup 3 --quick --no-cpu-baseline --per-op gpurun_out/sweep_$name.csv > gpurun_out/sweep_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_$name.log").read().strip().splitlines()[-1])
    print("$name: value %.0f nosync %.0f conv_ms %.3f plan_ms %.3f clocks %s" % (d["value"], d["config"]["value_no_host_sync"], d["roofline"]["step_share"]["conv_ms"], d["roofline"]["step_share"]["plan_ms"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$name: FAILED", e)
PY
}
run base YRE_X=0
run bn64 YRE_TC_BLOCK_N=64
run bn128 YRE_TC_BLOCK_N=128
run bn256 YRE_TC_BLOCK_N=256
run cta2off YRE_TC_CTA2=0
run cta2on YRE_TC_CTA2=1
run t384 YRE_TC_THREADS=384
run t512 YRE_TC_THREADS=512
run halo0 YRE_TC_HALO=0
run halo1 YRE_TC_HALO=1
run pipes1 YRE_TC_PIPES=1
run s64all YRE_TC_STAGE64=1
run s64none YRE_TC_STAGE64=0
run pairoff YRE_TC_HALO_PAIR=0
run split2 YRE_TC_SPLIT=2
run base2 YRE_X=0
cp /tmp/base.so $LIB
