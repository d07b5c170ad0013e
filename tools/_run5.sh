B="--headline-only --no-e2e --no-cpu --steps 3 --warmup 3"
NC="ncu --set full --clock-control none --import-source on -c 1 -f"
cap() { # name kernel-regex skip workload
  timeout 300 $NC -k regex:$2 -s $3 -o gpurun_out/r01d_prof_$1 python bench.py --workload $4 $B > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
}
timeout 300 python bench.py --headline-only --no-cpu > gpurun_out/s6_headline.json 2>gpurun_out/s6_headline.err; echo plain rc=$?
cap matmul4x4_f32 tile_kernel 4 matmul4x4_f32_16M
cap matmul3x3_f64_1M tile_kernel 4 matmul3x3_f64_1M
cap dmma128 dmma_tma 3 matmul128_f64_64K
cap dmma64 dmma_tma 3 matmul64_f64_256K
cap soa_matmul4x4_f32 soa_tile 3 matmul4x4_f32_soa_16M
cap inverse4_f64 tile_kernel 4 inverse4_f64_4M
cap prod1_8x8 rowprod 3 prod1_8x8_f64_1M
cap append gather2 3 append_3x3_3x1_f64_4M
cap tensorprod outer_kernel 3 tensorprod8x8_f64_1M
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_headline.csv python bench.py --headline-only --no-e2e --no-cpu --steps 50 --warmup 5 > gpurun_out/ncu_launches.log 2>&1; echo launches rc=$?
ls gpurun_out/r01d_*
