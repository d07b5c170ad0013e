set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s3_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s3_pytest.log
S=$(date +%s); timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo bench rc=$? secs=$(( $(date +%s) - S ))
B="--headline-only --no-e2e --no-cpu --steps 3 --warmup 3"
NC="ncu --set full --clock-control none --import-source on -c 1"
timeout 300 $NC -k regex:dmma_tma -s 3 -o gpurun_out/r01c_prof_dmma128_tma -f python bench.py --workload matmul128_f64_64K $B > gpurun_out/ncu_a.log 2>&1; echo rc=$?
timeout 300 $NC -k regex:dmma_tma -s 3 -o gpurun_out/r01c_prof_dmma64_tma -f python bench.py --workload matmul64_f64_256K $B > gpurun_out/ncu_b.log 2>&1; echo rc=$?
timeout 300 $NC -k regex:tile_kernel -s 4 -o gpurun_out/r01c_prof_matmul3x3_1M -f python bench.py --workload matmul3x3_f64_1M $B > gpurun_out/ncu_c.log 2>&1; echo rc=$?
timeout 300 $NC -k regex:gather2 -s 3 -o gpurun_out/r01c_prof_append -f python bench.py --workload append_3x3_3x1_f64_4M $B > gpurun_out/ncu_d.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
