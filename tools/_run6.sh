timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_entries.py -x -q -m gpu > gpurun_out/s7_pytest_multi.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/s7_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/s7_bench_n2.json 2> gpurun_out/s7_bench_n2.err; echo bench2 rc=$?
cut -c1-300 gpurun_out/s7_bench_n2.json
