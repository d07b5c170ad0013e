timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s8_pytest.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/s8_pytest.log
for w in matmul128_f32_64K dotsum4_f32_16M dotsum4_f32_soa_16M; do
timeout 300 python bench.py --workload $w --headline-only --no-e2e --no-cpu --steps 10 --warmup 3 2>gpurun_out/s8_$w.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', d['ms_per_step'], d['roofline'])"
done
