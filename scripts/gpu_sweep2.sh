#!/bin/bash
# Second per-layer sweep (same protocol as gpu_sweep.sh) over the knobs the first one left out.  usage: gpu_sweep2.sh TUNE.so [config]
mkdir -p gpurun_out
VAR=$1; CFG=${2:-2}
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
run pdl0 YRE_TC_PDL=0
run st3 YRE_TC_STAGES=3
run st4 YRE_TC_STAGES=4
run st5 YRE_TC_STAGES=5
run tg1 YRE_TC_TGROUPS=1
run tg2 YRE_TC_TGROUPS=2
run split2 YRE_TC_SPLIT=2
run pair64off YRE_TC_HALO_PAIR64=0
run direct YRE_TC_DIRECT_STORE=1
run s64all YRE_TC_STAGE64=1
run s64none YRE_TC_STAGE64=0
run cta2on YRE_TC_CTA2=1
run t512 YRE_TC_THREADS=512
run promo0 YRE_TC_L2PROMO_A=0
run promo256 YRE_TC_L2PROMO_A=3
run base2 YRE_X=0
cp /tmp/base.so $LIB
