"""Developer diagnostic for the tcgen05 3xTF32 product: error maps against float64 for structured inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_b200 as tb

def run(ctx, a, b, m, k, n, name):
    C = tb.matmul(ctx.from_numpy(a.astype(np.float32), (m, k)), ctx.from_numpy(b.astype(np.float32), (k, n))).to_numpy()
    want = np.einsum("bij,bjk->bik", a.astype(np.float64), b.astype(np.float64))
    err = np.abs(C - want)
    scale = np.einsum("bij,bjk->bik", np.abs(a), np.abs(b)) + 1e-30
    rel = err / scale
    print(f"{name}: max rel {rel.max():.3e}  mean rel {rel.mean():.3e}  nan {np.isnan(C).sum()}", flush=True)
    if rel.max() > 1e-5:
        bad = rel[0] > 1e-5
        print("  bad rows:", np.flatnonzero(bad.any(axis=1))[:16], " bad cols:", np.flatnonzero(bad.any(axis=0))[:16])
        print("  C[0,:4,:6]\n", C[0, :4, :6], "\n  want\n", want[0, :4, :6])

with tb.Context([0]) as ctx:
    rng = np.random.default_rng(0)
    m = k = n = 128
    eye = np.eye(128)[None]
    a = rng.standard_normal((1, m, k))
    run(ctx, a, eye, m, k, n, "A.I")
    run(ctx, eye, a, m, k, n, "I.B")
    idx = np.arange(128, dtype=np.float64)
    rowid = np.broadcast_to(idx[:, None], (1, 128, 128)).copy()      # A[i][j] = i
    colid = np.broadcast_to(idx[None, :], (1, 128, 128)).copy()      # B[i][j] = j
    run(ctx, rowid, eye, m, k, n, "rowid.I")
    run(ctx, eye, colid, m, k, n, "I.colid")
    run(ctx, a, rng.standard_normal((1, k, n)), m, k, n, "random")
    b5 = rng.standard_normal((5, 64, 256)); a5 = rng.standard_normal((5, 256, 64))
    run(ctx, a5, b5, 256, 64, 256, "256x64x256 batch 5")
    a9 = rng.standard_normal((300, 128, 128)); b9 = rng.standard_normal((300, 128, 128))
    run(ctx, a9, b9, 128, 128, 128, "batch 300")
