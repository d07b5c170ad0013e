for v in 0 1 2 3 4; do
  T5_MID8=$v timeout 200 python bench.py --workload prod1_8x8_f64_1M --headline-only --no-e2e --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('v=$v', d['ms_per_step'], d['roofline']['frac'], d['config']['l2'][:40])"
done
timeout 100 ./build/tune_mid | head -12
