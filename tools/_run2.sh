timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s4_pytest.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo bench rc=$?
timeout 300 ./build/tune_mid > gpurun_out/r01_tune_mid.csv 2>&1; echo tune rc=$?
cat gpurun_out/r01_tune_mid.csv
