mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_eqf.py -q -m gpu -k "lv_group_passes or equal_flow_reduced or text_loader or batch_driver_with_launch_groups" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2_sanitizer_memcheck.log | tail -5
