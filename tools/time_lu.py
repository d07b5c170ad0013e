"""Times det / inverse / solve of the CTA-per-matrix elimination kernels (9 <= n <= 128) on device-resident batches.
Prints ms, matrices/s, achieved GB/s (algorithmic bytes) and the unfused FP64/FP32 instruction rate in TFLOP/s
(multiply and subtract counted separately, as executed).  Usage (GPU box): python tools/time_lu.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb  # noqa: E402

SEED = 0x5EED000000000001


def timed(ctx, fn, reps=5):
    outs = [fn() for _ in range(reps)]          # as many live results as the timed loop holds: the allocator pool reaches its size
    for o in outs:
        for x in (o if isinstance(o, tuple) else (o,)):
            x.free()
    ctx.sync()
    ctx.timer_begin()
    outs = [fn() for _ in range(reps)]
    ms = ctx.timer_end() / reps
    for o in outs:
        for x in (o if isinstance(o, tuple) else (o,)):
            x.free()
    return ms


def main():
    with tb.Context([0]) as ctx:
        print("%-28s %9s %12s %9s %9s" % ("op", "ms", "matrices/s", "GB/s", "TFLOP/s"))
        for dt, es, tag in ((np.float64, 8, "f64"), (np.float32, 4, "f32")):
            for n in [int(x) for x in os.environ.get("LU_NS", "9,12,16,24,32,48,64,96,128").split(",")]:
                batch = max(256, min(1 << 20, (1 << 30) // (n * n * es)))
                A = ctx.synthetic((n, n), dt, batch, SEED, 5).plant_specials(1024, 65536)
                b = ctx.synthetic((n,), dt, batch, SEED + 1, 5)
                lu_flops = 2.0 * sum((n - k - 1) ** 2 for k in range(n))
                for name, fn, byts, flops in (
                        ("det", lambda: tb.det(A), (n * n + 1) * es, lu_flops),
                        ("solve (1 rhs)", lambda: tb.solve(A, b), (n * n + 2 * n) * es + 1, lu_flops + 2.0 * n * n),
                        ("inverse", lambda: tb.inverse(A), 2 * n * n * es + 1, lu_flops + 2.0 * n * n * n)):
                    ms = timed(ctx, fn)
                    print("%-28s %9.3f %12.3e %9.0f %9.2f" % (f"{name} {n}x{n} {tag} x{batch}", ms, batch / ms * 1e3,
                                                              byts * batch / ms / 1e6, flops * batch / ms / 1e9), flush=True)
                A.free(); b.free()


if __name__ == "__main__":
    main()
