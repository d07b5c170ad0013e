"""Developer survey: time every remaining public operation once (device-resident, CUDA events on the
library's stream) and print achieved GB/s against algorithmic bytes, to find untuned paths.
Each call allocates its result from the context's stream-ordered pool (warmed to steady size first), so the
figures are what a chained caller sees per operation.  Usage (GPU box): python tools/survey_ops.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb  # noqa: E402
from tensor_b200 import structure as st  # noqa: E402

SEED = 0x5EED000000000001


def timed(ctx, fn, reps=5):
    outs = [fn() for _ in range(reps)]      # as many live outputs as the timed loop: the pool reaches its steady size
    for o in outs:
        free(o)
    ctx.sync()
    ctx.timer_begin()
    outs = [fn() for _ in range(reps)]
    ms = ctx.timer_end() / reps
    for o in outs:
        free(o)
    return ms


def free(o):
    if isinstance(o, tuple):
        for x in o:
            free(x)
    elif hasattr(o, "free"):
        o.free()


def main():
    with tb.Context([0]) as ctx:
        rows = []

        def add(name, nbytes, fn):
            ms = min(timed(ctx, fn), timed(ctx, fn))      # best of two: the first pass can include the pool growing to a new high-water mark
            rows.append((name, ms, nbytes / ms / 1e6))
            print("%-44s %8.3f ms %8.0f GB/s" % rows[-1], flush=True)

        B = 1 << 22
        for dt, es in ((np.float64, 8), (np.float32, 4)):
            tag = "f64" if es == 8 else "f32"
            A3 = ctx.synthetic((3, 3), dt, B, SEED, 1)
            A4 = ctx.synthetic((4, 4), dt, B, SEED, 2)
            B4 = ctx.synthetic((4, 4), dt, B, SEED + 1, 2)
            v3 = ctx.synthetic((3, 1), dt, B, SEED, 3)
            A6 = ctx.synthetic((6, 6), dt, B // 4, SEED, 4)
            b6 = ctx.synthetic((6,), dt, B // 4, SEED + 1, 4)
            A8 = ctx.synthetic((8, 8), dt, B // 4, SEED, 5)
            T234 = ctx.synthetic((2, 3, 4), dt, B, SEED, 6)
            timed(ctx, lambda: tb.zip_with("+", A4, B4))        # grows the allocator pool once, untimed
            add(f"zip + 4x4 {tag}", B * 16 * es * 3, lambda: tb.zip_with("+", A4, B4))
            add(f"scale 4x4 {tag}", B * 16 * es * 2, lambda: tb.scale(dt(3.0), A4))
            add(f"tr 4x4 {tag}", B * (16 + 1) * es, lambda: tb.tr(A4))
            add(f"polyEval deg3 4x4 {tag}", B * 32 * es, lambda: tb.poly_eval(A4, [1.0, 2.0, 0.5, 0.25]))
            add(f"charPoly 4x4 {tag}", B * (16 + 5) * es, lambda: tb.char_poly(A4))
            add(f"minPoly 4x4 {tag}", B * (16 + 5) * es + B, lambda: tb.min_poly(A4))
            add(f"append 3x3|3 {tag}", B * 24 * es, lambda: st.append(2, A3, v3))
            add(f"split 4x4 -> 4x3,4x1 {tag}", B * 32 * es, lambda: st.split(2, 3, A4))
            add(f"at 4x4 column {tag}", B * (16 + 4) * es, lambda: st.at(A4, [2]))
            add(f"ta 4x4 row {tag}", B * (16 + 4) * es, lambda: st.ta(2, A4))
            add(f"slice 2x3x4 [_,2,_] {tag}", B * (24 + 8) * es, lambda: st.slice_(T234, [None, 2, None]))
            add(f"permute_axes 2x3x4 (2,0,1) {tag}", B * 48 * es, lambda: st.permute_axes(T234, [2, 0, 1]))
            add(f"transpose 2x3x4 {tag}", B * 48 * es, lambda: tb.transpose(T234))
            A5e = ctx.synthetic((5, 5), dt, B // 4, SEED, 14); b5e = ctx.synthetic((5,), dt, B // 4, SEED + 1, 14)
            add(f"det 5x5 {tag}", (B // 4) * 26 * es, lambda: tb.det(A5e))
            add(f"inverse 5x5 {tag}", (B // 4) * 50 * es, lambda: tb.inverse(A5e))
            add(f"solve 5x5 {tag}", (B // 4) * 35 * es, lambda: tb.solve(A5e, b5e))
            add(f"det 6x6 {tag}", (B // 4) * 37 * es, lambda: tb.det(A6))
            add(f"inverse 6x6 {tag}", (B // 4) * 72 * es, lambda: tb.inverse(A6))
            add(f"solve 6x6 {tag}", (B // 4) * 48 * es, lambda: tb.solve(A6, b6))
            add(f"det 8x8 {tag}", (B // 4) * 65 * es, lambda: tb.det(A8))
            b8 = ctx.synthetic((8,), dt, B // 4, SEED + 1, 15)
            add(f"solve 8x8 {tag}", (B // 4) * 80 * es, lambda: tb.solve(A8, b8))
            add(f"inverse 8x8 {tag}", (B // 4) * 128 * es, lambda: tb.inverse(A8))
            A45 = ctx.synthetic((4, 5), dt, B, SEED, 16)
            add(f"rowEchelonForm 4x4 {tag}", B * (32 * es + 1), lambda: tb.row_echelon_form(A4))
            add(f"rowEchelonForm 4x5 (triangularSolve) {tag}", B * (40 * es + 1), lambda: tb.row_echelon_form(A45, 4))
            A67 = ctx.synthetic((6, 7), dt, B // 4, SEED, 30)
            A79 = ctx.synthetic((7, 9), dt, B // 4, SEED, 31)
            add(f"rowEchelonForm 6x7 {tag}", (B // 4) * (84 * es + 1), lambda: tb.row_echelon_form(A67, 6))
            add(f"rowEchelonForm 7x9 (any-shape kernel) {tag}", (B // 4) * (126 * es + 1), lambda: tb.row_echelon_form(A79))
            A67.free(); A79.free()
            add(f"rowEchelonForm 8x8 {tag}", (B // 4) * (128 * es + 1), lambda: tb.row_echelon_form(A8))
            A5, B5 = ctx.synthetic((5, 5), dt, B, SEED, 7), ctx.synthetic((5, 5), dt, B, SEED, 8)
            add(f"matmul 5x5 {tag}", B * 75 * es, lambda: tb.matmul(A5, B5))
            add(f"matmul 4x4 (Double/Float) {tag}", B * 48 * es, lambda: tb.matmul(A4, B4))
            A32, B32 = ctx.synthetic((32, 32), dt, B // 64, SEED, 9), ctx.synthetic((32, 32), dt, B // 64, SEED, 10)
            add(f"matmul 32x32 {tag}", (B // 64) * 3072 * es, lambda: tb.matmul(A32, B32))
            # second sweep: shapes around the benchmarked ones (finds untuned dispatch corners)
            B8 = ctx.synthetic((8, 8), dt, B // 4, SEED + 1, 17)
            v8 = ctx.synthetic((8,), dt, B // 4, SEED + 1, 18)
            v64a, v64b = ctx.synthetic((64,), dt, B // 4, SEED, 19), ctx.synthetic((64,), dt, B // 4, SEED + 1, 19)
            v4a, v4b = ctx.synthetic((4,), dt, B, SEED, 20), ctx.synthetic((4,), dt, B, SEED + 1, 20)
            A2, B2 = ctx.synthetic((2, 2), dt, B, SEED, 21), ctx.synthetic((2, 2), dt, B, SEED + 1, 21)
            A16, B16 = ctx.synthetic((16, 16), dt, B // 16, SEED, 22), ctx.synthetic((16, 16), dt, B // 16, SEED + 1, 22)
            A12, B12 = ctx.synthetic((12, 12), dt, B // 16, SEED, 23), ctx.synthetic((12, 12), dt, B // 16, SEED + 1, 23)
            I8 = ctx.synthetic((8, 8), dt, B // 4, SEED + 2, 24)
            B44 = ctx.synthetic((4, 4), dt, B, SEED + 2, 25)
            add(f"prod 2: 8x8 : 8x8 -> scalar {tag}", (B // 4) * 129 * es, lambda: tb.prod(A8, B8, 2))
            add(f"dot 64 {tag}", (B // 4) * 129 * es, lambda: tb.dot(v64a, v64b))
            add(f"matvec 8x8 . 8 {tag}", (B // 4) * 80 * es, lambda: tb.matmul(A8, v8))
            add(f"matmul 8x8 {tag}", (B // 4) * 192 * es, lambda: tb.matmul(A8, B8))
            add(f"matmul 2x2 {tag}", B * 12 * es, lambda: tb.matmul(A2, B2))
            add(f"matmul 12x12 {tag}", (B // 16) * 432 * es, lambda: tb.matmul(A12, B12))
            add(f"matmul 16x16 {tag}", (B // 16) * 768 * es, lambda: tb.matmul(A16, B16))
            add(f"outer 4 (x) 4 {tag}", B * 24 * es, lambda: tb.prod(v4a, v4b, 0))
            add(f"outer 3x3 (x) 3x3 {tag}", B * 99 * es, lambda: tb.prod(A3, A3, 0))
            add(f"transpose 3x3 {tag}", B * 18 * es, lambda: tb.transpose(A3))
            add(f"transpose 8x8 {tag}", (B // 4) * 128 * es, lambda: tb.transpose(A8))
            add(f"transpose 16x16 {tag}", (B // 16) * 512 * es, lambda: tb.transpose(A16))
            A128 = ctx.synthetic((128, 128), dt, B // 1024, SEED, 26)
            A2040 = ctx.synthetic((20, 40), dt, B // 64, SEED, 27)
            add(f"transpose 128x128 {tag}", (B // 1024) * 32768 * es, lambda: tb.transpose(A128))
            add(f"transpose 20x40 {tag}", (B // 64) * 1600 * es, lambda: tb.transpose(A2040))
            A128.free(); A2040.free()
            T234s, A3s, v3s = (ctx.synthetic((2, 3, 4), dt, B, SEED, 6, tb.SOA), ctx.synthetic((3, 3), dt, B, SEED, 1, tb.SOA),
                               ctx.synthetic((3, 1), dt, B, SEED, 3, tb.SOA))
            A8s, B8s = ctx.synthetic((8, 8), dt, B // 4, SEED, 5, tb.SOA), ctx.synthetic((8, 8), dt, B // 4, SEED + 1, 17, tb.SOA)
            timed(ctx, lambda: st.append(2, A3s, v3s))          # (lets the allocator pool settle on the SoA result sizes)
            add(f"SoA transpose 2x3x4 {tag}", B * 48 * es, lambda: tb.transpose(T234s))
            add(f"SoA append 3x3|3 {tag}", B * 24 * es, lambda: st.append(2, A3s, v3s))
            # SoA: a slice copies 8 whole rows; the 16 rows it does not select are never read (AoS reads every sector of the item)
            add(f"SoA slice 2x3x4 [_,2,_] {tag}", B * (8 + 8) * es, lambda: st.slice_(T234s, [None, 2, None]))
            add(f"SoA matmul 8x8 (B in registers, A rows streamed) {tag}", (B // 4) * 192 * es, lambda: tb.matmul(A8s, B8s))
            for x in (T234s, A3s, v3s, A8s, B8s):
                x.free()
            b3s, b4s = ctx.synthetic((3,), dt, B, SEED + 1, 28), ctx.synthetic((4,), dt, B, SEED + 1, 29)
            for nn, AA, bb in ((3, A3, b3s), (4, A4, b4s)):
                add(f"det {nn}x{nn} {tag}", B * (nn * nn + 1) * es, lambda: tb.det(AA))
                add(f"inverse {nn}x{nn} {tag}", B * (2 * nn * nn * es + 1), lambda: tb.inverse(AA))
                add(f"solve {nn}x{nn} {tag}", B * ((nn * nn + 2 * nn) * es + 1), lambda: tb.solve(AA, bb))
            b3s.free(); b4s.free()
            add(f"det 2x2 {tag}", B * 5 * es, lambda: tb.det(A2))
            add(f"inverse 2x2 {tag}", B * (8 * es) + B, lambda: tb.inverse(A2))
            add(f"solve 4x4, 4 rhs {tag}", B * (48 * es) + B, lambda: tb.solve(A4, B44))
            add(f"solve 8x8, 8 rhs {tag}", (B // 4) * (192 * es) + B // 4, lambda: tb.solve(A8, I8))
            add(f"pure 4x4 {tag}", B * 16 * es, lambda: tb.pure(ctx, 1.5, (4, 4), dt, B))
            for x in (B8, v8, v64a, v64b, v4a, v4b, A2, B2, A16, B16, A12, B12, I8, B44):
                x.free()
            # round 2: eliminations beyond n = 8 (a CTA, or for Float 9..16 / 30..32 a half-warp / warp, per matrix), the SoA detour
            # of an AoS-only kernel, ragged and item-packed tensor-core products
            for nn in (16, 32, 64, 128):
                Bn = max(256, (B * 16) // (nn * nn))
                An = ctx.synthetic((nn, nn), dt, Bn, SEED, 40).plant_specials(1024, 65536)
                bn = ctx.synthetic((nn,), dt, Bn, SEED + 1, 40)
                add(f"det {nn}x{nn} (n > 8 kernels) {tag}", Bn * (nn * nn + 1) * es, lambda: tb.det(An))
                add(f"solve {nn}x{nn}, 1 rhs (n > 8 kernels) {tag}", Bn * ((nn * nn + 2 * nn) * es + 1), lambda: tb.solve(An, bn))
                add(f"inverse {nn}x{nn} (n > 8 kernels) {tag}", Bn * (2 * nn * nn * es + 1), lambda: tb.inverse(An))
                An.free(); bn.free()
            A32s, B32s = A32.relayout(tb.SOA), B32.relayout(tb.SOA)
            add(f"SoA matmul 32x32 (detour: relayout in, AoS kernel, relayout out) {tag}", (B // 64) * 3072 * es, lambda: tb.matmul(A32s, B32s))
            A32s.free(); B32s.free()
            A96, B96 = ctx.synthetic((96, 96), dt, B // 64, SEED, 41), ctx.synthetic((96, 96), dt, B // 64, SEED + 1, 41)
            add(f"matmul 96x96 ({'ragged tcgen05 tiles' if es == 4 else 'DMMA 96 tile'}) {tag}", (B // 64) * 3 * 9216 * es, lambda: tb.matmul(A96, B96))
            A96.free(); B96.free()
            for nn in (24, 48, 64):
                Bn = max(256, (B * 16 * 3) // (3 * nn * nn))
                An, Bm = ctx.synthetic((nn, nn), dt, Bn, SEED, 42), ctx.synthetic((nn, nn), dt, Bn, SEED + 1, 42)
                cls = "exact" if tb.matmul_is_exact(tb.device.DTYPE_OF[np.dtype(dt)], nn, nn, nn) else "tensor cores"
                add(f"matmul {nn}x{nn} ({cls}) {tag}", Bn * 3 * nn * nn * es, lambda: tb.matmul(An, Bm))
                An.free(); Bm.free()
            for x in (A3, A4, B4, v3, A6, b6, A8, T234, A5, B5, A32, B32):
                x.free()


if __name__ == "__main__":
    main()
