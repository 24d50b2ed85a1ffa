timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_round2.py tests/test_gpu_model.py -q -p no:cacheprovider -k "nms or config5 or fused_scale or concurrent or async or class_count or end_to_end" 2>&1 | tail -6
python scripts/nms_diag.py 2>&1 | tail -14
