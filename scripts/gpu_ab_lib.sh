#!/bin/bash
# A/B of two builds of libyre.so on ONE box: swaps the library file between bench runs (the box's repo copy is scratch).
# usage: gpu_ab_lib.sh VARIANT.so [config]
mkdir -p gpurun_out
VAR=$1; CFG=${2:-2}
LIB=yolo-re_b200/yolo_b200/libyre.so
cp $LIB /tmp/base.so
for i in 1 2; do
  for which in base var; do
    if [ $which = base ]; then cp /tmp/base.so $LIB; else cp $VAR $LIB; fi
    timeout 600 python bench.py --config $CFG --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/abl_${which}_$i.log 2>&1
    python - <<PY
import json
d=json.loads(open("gpurun_out/abl_${which}_$i.log").read().strip().splitlines()[-1])
print("$which run $i: value %.0f e2e %.0f nosync %.0f clocks %s" % (d["value"], d["e2e"]["value"], d["config"]["value_no_host_sync"], d["clocks"]["sm_mhz"]))
PY
  done
done
cp /tmp/base.so $LIB
