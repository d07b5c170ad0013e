"""Developer sweep: square .*. for n = 2..24 and a few rectangular shapes, every element type (device-resident,
~256 MB per operand), achieved GB/s against algorithmic bytes - finds dispatch holes between the specialised kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb
from survey_ops import timed, SEED
with tb.Context([0]) as ctx:
    shapes = [(n, n, n) for n in list(range(2, 21)) + [24, 32, 48]] + [(3, 4, 1), (4, 3, 2), (2, 8, 2), (8, 4, 8), (16, 16, 1), (1, 16, 16), (4, 16, 4), (32, 32, 1), (9, 9, 1), (12, 12, 1), (24, 24, 1)]
    for dt, es in ((np.float64, 8), (np.float32, 4), (np.int64, 8)):
        for (m, k, n) in shapes:
            B = max(256, (256 << 20) // ((m * k + k * n + m * n) * es))
            A, Bm = ctx.synthetic((m, k), dt, B, SEED, 60), ctx.synthetic((k, n), dt, B, SEED + 1, 61)
            ms = timed(ctx, lambda: tb.matmul(A, Bm))
            print("matmul %2dx%2dx%2d %-8s %8.3f ms %6.0f GB/s" % (m, k, n, dt.__name__, ms, B * (m * k + k * n + m * n) * es / ms / 1e6), flush=True)
            A.free(); Bm.free()
