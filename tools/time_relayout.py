"""Developer timing: t5_relayout in both directions over record sizes either side of the 64-element staging limit, and
the SoA 8x8 contraction, the structural gathers, long dots, the any-shape echelon kernel and the SoA outer product.  (The
environment variables some labels mention selected kernel variants while these were tuned - the A/B files under profiles/
r02f .. r02q - and no longer exist in the library.)  Device-resident, CUDA events on
the library's stream, GB/s against read + written bytes.  Usage (GPU box): python tools/time_relayout.py [prod | struct | dot | echelon | outer]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb  # noqa: E402
from tensor_b200 import structure as st  # noqa: E402

SEED = 0x5EED000000000001


def free(o):
    if isinstance(o, tuple):
        for x in o:
            free(x)
    else:
        o.free()


def timed(ctx, fn, reps=10):
    outs = [fn() for _ in range(reps)]
    for o in outs:
        free(o)
    ctx.sync()
    best = 1e9
    for _ in range(3):
        ctx.timer_begin()
        outs = [fn() for _ in range(reps)]
        best = min(best, ctx.timer_end() / reps)
        for o in outs:
            free(o)
    return best


def structural(ctx):
    """The table-driven AoS gathers (T5_NO_STAGED_GATHER=1: the direct one-element-per-lane kernels)."""
    B = 1 << 22
    for dt, es in ((np.float32, 4), (np.float64, 8)):
        tag = "f32" if es == 4 else "f64"
        A3, v3 = ctx.synthetic((3, 3), dt, B, SEED, 1), ctx.synthetic((3, 1), dt, B, SEED, 3)
        A4 = ctx.synthetic((4, 4), dt, B, SEED, 2)
        T234 = ctx.synthetic((2, 3, 4), dt, B, SEED, 6)
        A8 = ctx.synthetic((8, 8), dt, B // 4, SEED, 5)
        for name, nbytes, fn in (("append 3x3|3", B * 24 * es, lambda: st.append(2, A3, v3)),
                                 ("split 4x4 -> 4x3,4x1", B * 32 * es, lambda: st.split(2, 3, A4)),
                                 ("at 4x4 column", B * 20 * es, lambda: st.at(A4, [2])),
                                 ("ta 4x4 row", B * 20 * es, lambda: st.ta(2, A4)),
                                 ("slice 2x3x4 [_,2,_]", B * 32 * es, lambda: st.slice_(T234, [None, 2, None])),
                                 ("permute_axes 2x3x4 (2,0,1)", B * 48 * es, lambda: st.permute_axes(T234, [2, 0, 1])),
                                 ("transpose 2x3x4", B * 48 * es, lambda: tb.transpose(T234)),
                                 ("transpose 8x8", (B // 4) * 128 * es, lambda: tb.transpose(A8)),
                                 ("split 8x8 -> 8x5,8x3", (B // 4) * 128 * es, lambda: st.split(2, 5, A8))):
            ms = timed(ctx, fn)
            print("%-28s %s %7.3f ms %6.0f GB/s%s" % (name, tag, ms, nbytes / ms / 1e6,
                                                      "  (direct kernels)" if os.environ.get("T5_NO_STAGED_GATHER") else ""), flush=True)
        for x in (A3, v3, A4, T234, A8):
            x.free()


def main():
    with tb.Context([0]) as ctx:
        if "struct" in sys.argv[1:]:
            return structural(ctx)
        if "staged" in sys.argv[1:]:
            for dt, es in ((np.float32, 4), (np.float64, 8), (np.int64, 8)):
                for (m, k, n) in ((21, 21, 21), (23, 23, 23), (22, 22, 22), (13, 7, 19), (5, 24, 9), (23, 5, 6), (9, 9, 5)):
                    batch = max(256, (256 << 20) // ((m * k + k * n + m * n) * es))
                    A, B = ctx.synthetic((m, k), dt, batch, SEED, 1), ctx.synthetic((k, n), dt, batch, SEED + 1, 1)
                    ms = timed(ctx, lambda: tb.matmul(A, B))
                    print("matmul %2dx%2dx%2d %-8s x %8d %7.3f ms %6.0f GB/s%s" % (m, k, n, np.dtype(dt).name, batch, ms, batch * (m * k + k * n + m * n) * es / ms / 1e6,
                          "  (one element per lane)" if os.environ.get("T5_STAGED_SCALAR") else ""), flush=True)
                    A.free(); B.free()
            return
        if "outer" in sys.argv[1:]:
            for dt, es in ((np.float64, 8), (np.float32, 4)):
                batch = 1 << 18
                A = ctx.synthetic((8, 8), dt, batch, SEED, 5, tb.SOA)
                B = ctx.synthetic((8, 8), dt, batch, SEED + 1, 17, tb.SOA)
                ms = timed(ctx, lambda: tb.tensor_product(A, B), reps=4)
                print("SoA 8x8 (x) 8x8 %-8s x %d  T5_SOA_OUTER_THREADS=%s: %7.3f ms %6.0f GB/s" % (np.dtype(dt).name, batch, os.environ.get("T5_SOA_OUTER_THREADS", "64"),
                                                                                              ms, batch * (4096 + 128) * es / ms / 1e6), flush=True)
                A.free(); B.free()
            return
        if "echelon" in sys.argv[1:]:
            for dt, es in ((np.float64, 8), (np.float32, 4)):
                for (m, n), batch in (((7, 9), 1 << 20), ((5, 7), 1 << 20), ((8, 12), 1 << 19), ((3, 7), 1 << 21), ((8, 16), 1 << 19)):
                    a = ctx.synthetic((m, n), dt, batch, SEED, 9)
                    ms = timed(ctx, lambda: tb.row_echelon_form(a))
                    print("rowEchelonForm %dx%-2d %-8s x %8d %7.3f ms %6.0f GB/s" % (m, n, np.dtype(dt).name, batch, ms, batch * (2 * m * n * es + 1) / ms / 1e6), flush=True)
                    a.free()
            return
        if "dot" in sys.argv[1:]:
            for dt, es in ((np.float32, 4), (np.float64, 8), (np.int64, 8)):
                for n, batch in ((64, 1 << 20), (16, 1 << 22), (256, 1 << 18), (100, 1 << 20)):
                    x, y = ctx.synthetic((n,), dt, batch, SEED, 7), ctx.synthetic((n,), dt, batch, SEED + 1, 7)
                    ms = timed(ctx, lambda: tb.dot(x, y))
                    print("dot %4d %-8s x %8d %7.3f ms %6.0f GB/s%s" % (n, np.dtype(dt).name, batch, ms, batch * (2 * n + 1) * es / ms / 1e6,
                                                                       "  (one element per lane)" if os.environ.get("T5_NO_LONGDOT_ROWS") else ""), flush=True)
                    x.free(); y.free()
            return
        for dt, es in ((np.float32, 4), (np.float64, 8)) if "prod" not in sys.argv[1:] else ():
            tag = "f32" if es == 4 else "f64"
            for shape, batch in (((4, 4), 1 << 24), ((3, 3), 1 << 24), ((2, 3), 1 << 24), ((8, 8), 1 << 22), ((32, 32), 1 << 18),
                                 ((128, 128), 1 << 14), ((10, 13), 1 << 21)):
                A = ctx.synthetic(shape, dt, batch, SEED, 1)
                S = A.relayout(tb.SOA)
                nbytes = 2 * batch * int(np.prod(shape)) * es
                ms_to = timed(ctx, lambda: A.relayout(tb.SOA))
                ms_from = timed(ctx, lambda: S.relayout(tb.AOS))
                print("relayout %-8s %s x %8d   AoS->SoA %7.3f ms %6.0f GB/s   SoA->AoS %7.3f ms %6.0f GB/s"
                      % ("x".join(map(str, shape)), tag, batch, ms_to, nbytes / ms_to / 1e6, ms_from, nbytes / ms_from / 1e6), flush=True)
                A.free(); S.free()
        for dt, es in ((np.float64, 8), (np.int64, 8), (np.float32, 4)):
            batch = 1 << 20
            A = ctx.synthetic((8, 8), dt, batch, SEED, 5, tb.SOA)
            B = ctx.synthetic((8, 8), dt, batch, SEED + 1, 17, tb.SOA)
            ms = timed(ctx, lambda: tb.matmul(A, B))
            print("SoA matmul 8x8 %s variant %s: %7.3f ms %6.0f GB/s" % (np.dtype(dt).name, os.environ.get("T5_SOA_PROD_VARIANT", "default"),
                                                                    ms, batch * 192 * es / ms / 1e6), flush=True)
            A.free(); B.free()


if __name__ == "__main__":
    main()
