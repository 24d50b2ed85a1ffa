#!/bin/bash
# One GPU-box session: diagnostics, parity tests, bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tc_diag"; timeout 900 python scripts/tc_diag.py all 2>&1 | tee gpurun_out/tc_diag.log
echo "== pytest fp32/nms/decode"; timeout 1500 python -m pytest tests -m gpu -q -x -k "not bf16" 2>&1 | tail -40 | tee gpurun_out/pytest_fp32.log
echo "== pytest bf16"; timeout 900 python -m pytest tests -m gpu -q -k "bf16" 2>&1 | tail -60 | tee gpurun_out/pytest_bf16.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
