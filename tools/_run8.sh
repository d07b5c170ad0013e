export PYTHONUNBUFFERED=1
K="test_matmul_bit_exact or test_det_inverse_solve_bit_exact or test_dmma_matmul_within_tolerance or test_exact_blocked or test_append_split or test_transpose_bit_exact or test_dot_sum or test_row_per_thread or test_matmul_ragged"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all python -m pytest tests/test_gpu_parity.py tests/test_gpu_structural.py -x -q -m gpu -k "$K" > gpurun_out/s9_memcheck.log 2>&1; echo memcheck rc=$?
tail -8 gpurun_out/s9_memcheck.log
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/s9_memcheck.log
K2="test_matmul_ragged or test_row_per_thread or test_exact_blocked_matmul_specials"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K2" > gpurun_out/s9_racecheck.log 2>&1; echo racecheck rc=$?
tail -8 gpurun_out/s9_racecheck.log
