timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s5_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s5_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo bench rc=$?
