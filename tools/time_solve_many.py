"""Developer timing of `solve` with matrix right-hand sides for 5 <= n <= 8 (device-resident, 1Mi items)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb
from survey_ops import timed, SEED
with tb.Context([0]) as ctx:
    B = 1 << 20
    for dt, es in ((np.float64, 8), (np.float32, 4)):
        for n in (5, 6, 7, 8):
            A = ctx.synthetic((n, n), dt, B, SEED, 40 + n)
            for r in (2, 4, n):
                R = ctx.synthetic((n, r), dt, B, SEED + 1, 50 + r)
                ms = timed(ctx, lambda: tb.solve(A, R))
                print("solve %dx%d nrhs=%d %s %.3f ms %.0f GB/s" % (n, n, r, dt.__name__, ms, B * ((n * n + 2 * n * r) * es + 1) / ms / 1e6), flush=True)
                R.free()
            A.free()
