"""Developer timing: Float .*. on small / ragged shapes (tcgen05 3xTF32: four 32 x 32 items per 128 tile) - GB/s against
algorithmic bytes.  Usage: time_tf32_small.py [m,k,n ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb
from survey_ops import timed, SEED
shapes = [(16, 16, 16), (24, 24, 24), (32, 32, 32), (32, 16, 32), (25, 44, 32), (32, 64, 32), (48, 48, 48), (64, 64, 64), (96, 96, 96), (128, 128, 128), (64, 40, 64)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
with tb.Context([0]) as ctx:
    for (m, k, n) in shapes:
        B = max(64, (1 << 29) // ((m * k + k * n + m * n) * 4))
        A, Bm = ctx.synthetic((m, k), np.float32, B, SEED, 60), ctx.synthetic((k, n), np.float32, B, SEED + 1, 61)
        ms = timed(ctx, lambda: tb.matmul(A, Bm))
        print("matmul %3dx%3dx%3d f32 %s %8.3f ms %6.0f GB/s %6.2f TFLOP/s" % (m, k, n, "exact" if tb.matmul_is_exact(tb.F32, m, k, n) else "tcgen05",
              ms, B * (m * k + k * n + m * n) * 4 / ms / 1e6, 2.0 * B * m * k * n / ms / 1e9), flush=True)
        A.free(); Bm.free()
