"""Developer timing: Double .*. on shapes off the 64 / 128 tile grid (DMMA + TMA with zero-filled boxes), small items
packed four to a tile, and their neighbours on the grid; GB/s against algorithmic bytes and TFLOP/s.
Usage: time_dmma_ragged.py [m,k,n ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb
from survey_ops import timed, SEED
shapes = [(64, 64, 64), (96, 96, 96), (128, 128, 128), (100, 64, 104), (48, 32, 56), (72, 72, 72), (80, 80, 80),
          (112, 112, 112), (160, 160, 160), (192, 192, 192), (65, 128, 64), (40, 40, 40), (56, 56, 56)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
with tb.Context([0]) as ctx:
    for (m, k, n) in shapes:
        B = max(64, (1 << 30) // ((m * k + k * n + m * n) * 8))
        A, Bm = ctx.synthetic((m, k), np.float64, B, SEED, 60), ctx.synthetic((k, n), np.float64, B, SEED + 1, 61)
        ms = timed(ctx, lambda: tb.matmul(A, Bm))
        print("matmul %3dx%3dx%3d f64 %8.3f ms %6.0f GB/s %6.2f TFLOP/s" % (m, k, n, ms, B * (m * k + k * n + m * n) * 8 / ms / 1e6,
              2.0 * B * m * k * n / ms / 1e9), flush=True)
        A.free(); Bm.free()
