#!/bin/bash
# Per-layer sweep of the conv_tc configuration knobs (tuning build, env overrides) on ONE box: every variant writes a per-op
# CSV; scripts/sweep_table.py merges them into one table (best variant per layer).  usage: gpu_sweep.sh TUNE.so [config]
mkdir -p gpurun_out
VAR=$1; CFG=${2:-2}
LIB=yolo-re_b200/yolo_b200/libyre.so
cp $LIB /tmp/base.so; cp $VAR $LIB
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
run slack125 YRE_TC_HALO_SLACK=125
run halo0 YRE_TC_HALO=0
run halo1 YRE_TC_HALO=1
run pipes1 YRE_TC_PIPES=1
run s64off YRE_TC_STAGE64=0
run tps1 YRE_TC_HALO_TPS=1
run pairoff YRE_TC_HALO_PAIR=0
run bn128cta2 YRE_TC_BLOCK_N=128 YRE_TC_CTA2=1
run base2 YRE_X=0
cp /tmp/base.so $LIB
