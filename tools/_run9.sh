timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size" > gpurun_out/s10_pytest.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/s10_pytest.log
S=$(date +%s); timeout 900 python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; echo bench rc=$? secs=$(( $(date +%s) - S ))
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s10_bench_ref.json 2> gpurun_out/s10_bench_ref.err; echo ref rc=$? secs=$(( $(date +%s) - S ))
