#!/bin/bash
# compute-sanitizer passes over the kernel unit tests at small shapes: memcheck, racecheck and
# NOTE (round 2): the GPU pool refuses to run compute-sanitizer (profiles/r02_sanitizer_note.txt); tests/test_gpu_guard.py is the
# guard-band substitute for the memcheck pass.  The script is kept for pools where the tool is available.
# synccheck for the tcgen05 / TMA conv family, the bandwidth kernels and NMS.  Run on the GPU box:
#     bash scripts/gpu_sanitize.sh            # writes gpurun_out/sanitizer_<tool>.txt (+ a one-line verdict per tool)
# The summaries kept under profiles/r02_sanitizer_*.txt are the tail of these logs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
# small shapes only: the sanitizer slows kernels down 10-100x and every mbarrier wait has a 2 s watchdog
SEL='test_conv and bf16 or test_block and bf16 or test_stem_conv or test_upsample or test_decode_kernel or nms_bit_exact_vs_reference_fixture or test_detect_head'
SEL2='test_conv_cta_pair or test_stem_u8 or test_nms_fused'
for tool in memcheck racecheck synccheck; do
  out=gpurun_out/sanitizer_${tool}.txt
  echo "== $tool ==" > "$out"
  timeout 1500 $SAN --tool $tool --print-limit 20 --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$SEL" -p no:cacheprovider >> "$out" 2>&1
  rc1=$?
  timeout 1500 $SAN --tool $tool --print-limit 20 --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "$SEL2" -p no:cacheprovider >> "$out" 2>&1
  rc2=$?
  echo "[$tool] exit codes: ops=$rc1 round2=$rc2 (0 = tests passed and the sanitizer reported nothing; 77 = sanitizer errors; 124 = timeout)" | tee -a "$out"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$out" | tail -6
done
