#!/bin/bash
# Per-op timing of the conv_tc configuration variants (env overrides) on one box, for picking defaults.
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-op gpurun_out/per_op_$name.csv > gpurun_out/bench_$name.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$name.log") if x.startswith("{")]
j=json.loads(l[-1]) if l else None
print("$name", j and (round(j["value"]), j["stage_ms_per_step"]))
PY
}
run A YRE_X=0
run B YRE_TC_THREADS=384 YRE_TC_TGROUPS=1 YRE_TC_SPLIT=2
run C YRE_TC_THREADS=384
run D YRE_TC_THREADS=512
run A2 YRE_X=0
