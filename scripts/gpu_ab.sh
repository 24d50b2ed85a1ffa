#!/bin/bash
# A/B on ONE box (box-to-box clock spread is larger than a 2 % change): bench --quick with an environment switch off / on,
# alternating twice.  usage: gpu_ab.sh VAR [config]
mkdir -p gpurun_out
V=$1; CFG=${2:-2}
for i in 1 2; do
  for val in 0 1; do
    env $V=$val timeout 600 python bench.py --config $CFG --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/ab_${V}_${val}_$i.log 2>&1
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${V}_${val}_$i.log").read().strip().splitlines()[-1])
print("$V=$val run $i: value %.0f e2e %.0f nosync %.0f conv_ms %.3f plan_ms %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["config"]["value_no_host_sync"], d["roofline"]["step_share"]["conv_ms"], d["roofline"]["step_share"]["plan_ms"], d["clocks"]["sm_mhz"]))
PY
  done
done
