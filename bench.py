#!/usr/bin/env python
"""bench.py — throughput of the batched fixed-dimension linear-algebra hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of one operation over one synthetic batch.  Headline workload
(BASELINE.json configs[1]): 16 Mi pairs of 4x4 Float matrices through `.*.`.
Rank 0 prints ONE JSON line (see README/DESIGN for the keys).  Weak scaling: every
rank owns a full batch; no collective is on the data path.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED_A, SEED_B = 0x5EED000000000001, 0x5EED000000000002

# name -> description of the step; bytes = ALGORITHMIC bytes per item
WORKLOADS = {
    "matmul4x4_f32_16M": dict(op="matmul", dtype="f32", m=4, k=4, n=4, batch=1 << 24, bytes=192, bound="hbm",
                              desc="16Mi pairs of 4x4 Float matrices, .*. (BASELINE configs[1])"),
    "matmul3x3_f64_1M": dict(op="matmul", dtype="f64", m=3, k=3, n=3, batch=1 << 20, bytes=216, bound="hbm",
                             desc="1Mi pairs of 3x3 Double matrices, .*. (configs[0])"),
    "matmul3x3_i64_1M": dict(op="matmul", dtype="i64", m=3, k=3, n=3, batch=1 << 20, bytes=216, bound="hbm",
                             desc="1Mi pairs of 3x3 Int64 matrices, .*. (configs[0], bit-exact variant)"),
    "matmul3x3_f64_64M": dict(op="matmul", dtype="f64", m=3, k=3, n=3, batch=1 << 26, bytes=216, bound="hbm",
                              desc="64Mi pairs of 3x3 Double matrices, .*. (launch-amortised size, SURVEY H6)"),
    "transpose4x4_f32_16M": dict(op="transpose", dtype="f32", dims=(4, 4), batch=1 << 24, bytes=128, bound="hbm",
                                 desc="16Mi 4x4 Float transposes (configs[1])"),
    "matvec4_f32_16M": dict(op="matmul", dtype="f32", m=4, k=4, n=1, batch=1 << 24, bytes=96, bound="hbm",
                            desc="16Mi 4x4 . 4-vector Float (configs[1])"),
    "dot4_f32_16M": dict(op="dot", dtype="f32", n=4, batch=1 << 24, bytes=36, bound="hbm",
                         desc="16Mi 4-vector Float dot (configs[1])"),
    "det3_f64_4M": dict(op="det", dtype="f64", n=3, batch=1 << 22, bytes=80, bound="hbm", desc="4Mi 3x3 Double det (configs[2])"),
    "det4_f64_4M": dict(op="det", dtype="f64", n=4, batch=1 << 22, bytes=136, bound="hbm", desc="4Mi 4x4 Double det (configs[2])"),
    "inverse3_f64_4M": dict(op="inverse", dtype="f64", n=3, batch=1 << 22, bytes=145, bound="hbm", desc="4Mi 3x3 Double inverse (configs[2])"),
    "inverse4_f64_4M": dict(op="inverse", dtype="f64", n=4, batch=1 << 22, bytes=257, bound="hbm", desc="4Mi 4x4 Double inverse (configs[2])"),
    "solve3_f64_4M": dict(op="solve", dtype="f64", n=3, batch=1 << 22, bytes=121, bound="hbm", desc="4Mi 3x3 Double solve (configs[2])"),
    "solve4_f64_4M": dict(op="solve", dtype="f64", n=4, batch=1 << 22, bytes=193, bound="hbm", desc="4Mi 4x4 Double solve (configs[2])"),
    "tensorprod8x8_f64_1M": dict(op="prod", dtype="f64", dims_a=(8, 8), dims_b=(8, 8), nc=0, batch=1 << 20, bytes=33792, bound="hbm",
                                 desc="1Mi 8x8 (x) 8x8 -> 8^4 Double tensor products (configs[3])"),
    "prod1_8x8_f64_1M": dict(op="prod", dtype="f64", dims_a=(8, 8), dims_b=(8, 8), nc=1, batch=1 << 20, bytes=1536, bound="hbm",
                             desc="1Mi 8x8 . 8x8 Double contractions (configs[3], n=1)"),
    "matmul4x4_f32_soa_16M": dict(op="matmul", dtype="f32", m=4, k=4, n=4, batch=1 << 24, bytes=192, bound="hbm", layout="soa",
                                  desc="16Mi pairs of 4x4 Float matrices, .*., SoA device layout"),
    "inverse4_f64_soa_4M": dict(op="inverse", dtype="f64", n=4, batch=1 << 22, bytes=257, bound="hbm", layout="soa",
                                desc="4Mi 4x4 Double inverse, SoA device layout"),
    "append_3x3_3x1_f64_4M": dict(op="append", dtype="f64", dims_a=(3, 3), dims_b=(3, 1), axis=2, batch=1 << 22, bytes=192, bound="hbm",
                                  desc="4Mi augmented matrices [A | b] via append (structural gather)"),
    "transpose8x8x8_f64_1M": dict(op="transpose", dtype="f64", dims=(8, 8, 8), batch=1 << 20, bytes=8192, bound="hbm",
                                  desc="1Mi rank-3 8x8x8 Double transposes (any-rank gather)"),
    "matmul128_f64_64K": dict(op="matmul", dtype="f64", m=128, k=128, n=128, batch=1 << 16, bytes=393216, flops=4194304, bound="tensor",
                              desc="64Ki 128x128 Double .*. on FP64 tensor cores (configs[4])"),
}
HEADLINE = "matmul4x4_f32_16M"
EXTRAS = ["matmul3x3_f64_1M", "matmul3x3_i64_1M", "matmul3x3_f64_64M", "transpose4x4_f32_16M", "matvec4_f32_16M",
          "dot4_f32_16M", "det3_f64_4M", "det4_f64_4M", "inverse3_f64_4M", "inverse4_f64_4M", "solve3_f64_4M",
          "solve4_f64_4M", "tensorprod8x8_f64_1M", "prod1_8x8_f64_1M", "matmul4x4_f32_soa_16M", "inverse4_f64_soa_4M",
          "append_3x3_3x1_f64_4M", "transpose8x8x8_f64_1M"]
NPD = {"f64": np.float64, "f32": np.float32, "i64": np.int64}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            j = json.load(f)
        out = {"hbm_gbs": float(j["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        out = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    out["fp64_tflops"] = None
    try:  # cuBLAS DGEMM 8192^3 measured by tools/measure_fp64_peak.py on this pool's B200
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
            out["fp64_tflops"] = float(json.load(f)["fp64_tflops"])
    except Exception:
        pass
    return out


def load_traffic(name):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


# ---- clocks sampled during the timed region (NVML from a thread) ----------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- one workload on the device ---------------------------------------------------------------
class Step:
    """Allocates the operands of a workload in HBM and exposes run() = one kernel launch."""

    def __init__(self, ctx, tb, w, first_item=0):
        self.ctx, self.tb, self.w = ctx, tb, w
        L = tb._lib.lib()
        dt = NPD[w["dtype"]]
        B = w["batch"]
        op = w["op"]
        lay = tb.SOA if w.get("layout") == "soa" else tb.AOS
        syn = lambda shape, seed, opid: ctx.synthetic(shape, dt, B, seed, opid, lay)  # noqa: E731
        empty = lambda shape, d=dt: ctx.empty(shape, d, B, lay)  # noqa: E731
        h = ctx._h
        code = tb.device.DTYPE_OF[np.dtype(dt)]
        if op == "matmul":
            m, k, n = w["m"], w["k"], w["n"]
            self.a, self.b = syn((m, k), SEED_A, 1), syn((k, n), SEED_B, 1)
            self.c = empty((m, n))
            self.run = lambda: L.t5_matmul(h, code, m, k, n, self.a._h, self.b._h, self.c._h)
        elif op == "dot":
            n = w["n"]
            self.a, self.b = syn((n,), SEED_A, 2), syn((n,), SEED_B, 2)
            self.c = empty(())
            self.run = lambda: L.t5_dot(h, code, n, self.a._h, self.b._h, self.c._h)
        elif op == "transpose":
            dims = w["dims"]
            self.a = syn(dims, SEED_A, 3)
            self.c = empty(tuple(reversed(dims)))
            da = tb._lib.u64_array(dims)
            self.run = lambda: L.t5_transpose(h, code, len(dims), da, self.a._h, self.c._h)
        elif op == "prod":
            da, db, nc = w["dims_a"], w["dims_b"], w["nc"]
            self.a, self.b = syn(da, SEED_A, 4), syn(db, SEED_B, 4)
            self.c = empty(tuple(da[: len(da) - nc]) + tuple(db[nc:]))
            ua, ub = tb._lib.u64_array(da), tb._lib.u64_array(db)
            self.run = lambda: L.t5_prod(h, code, len(da), ua, len(db), ub, nc, self.a._h, self.b._h, self.c._h)
        elif op == "append":
            da, db, ax = w["dims_a"], w["dims_b"], w["axis"]
            self.a, self.b = syn(da, SEED_A, 6), syn(db, SEED_B, 6)
            self.c = empty(tuple(x + y if k == ax - 1 else x for k, (x, y) in enumerate(zip(da, db))))
            ua, ub = tb._lib.u64_array(da), tb._lib.u64_array(db)
            self.run = lambda: L.t5_append(h, code, ax, len(da), ua, ub, self.a._h, self.b._h, self.c._h)
        elif op in ("det", "inverse", "solve"):
            n = w["n"]
            self.a = syn((n, n), SEED_A, 5).plant_specials(1024, 65536)
            if op == "det":
                self.c = empty(())
                self.run = lambda: L.t5_det(h, code, n, self.a._h, self.c._h)
            elif op == "inverse":
                self.c, self.s = empty((n, n)), empty((), np.uint8)
                self.run = lambda: L.t5_inverse(h, code, n, self.a._h, self.c._h, self.s._h)
            else:
                self.b = syn((n,), SEED_B, 5)
                self.c, self.s = empty((n,)), empty((), np.uint8)
                self.run = lambda: L.t5_solve(h, code, n, 1, self.a._h, self.b._h, self.c._h, self.s._h)
        else:
            raise ValueError(op)

    def free(self):
        for nm in ("a", "b", "c", "s"):
            if hasattr(self, nm):
                getattr(self, nm).free()


def check_rc(rc, ctx, tb):
    if rc != 0:
        tb.device.check(rc, ctx)


class StepRing:
    """Several independent operand sets of one workload, used round-robin: together they are
    several times larger than the 126 MB L2, so every timed launch streams from HBM (the
    "inputs larger than L2" rule) without a dirty-line flush perturbing small workloads."""

    def __init__(self, ctx, tb, w, min_bytes=640 << 20, max_sets=8):
        per = w["bytes"] * w["batch"]
        n = 1 if per >= min_bytes else min(max_sets, -(-min_bytes // per))
        self.sets = [Step(ctx, tb, w) for _ in range(n)]
        self.i = 0

    def run(self):
        rc = self.sets[self.i].run()
        self.i = (self.i + 1) % len(self.sets)
        return rc

    @property
    def first(self):
        return self.sets[0]

    def free(self):
        for s in self.sets:
            s.free()


def time_steps(ctx, tb, ring, steps, warmup):
    """Total device milliseconds of `steps` back-to-back launches (CUDA events on the launch stream)."""
    for _ in range(warmup):
        check_rc(ring.run(), ctx, tb)
    ctx.sync()
    ctx.timer_begin()
    for _ in range(steps):
        check_rc(ring.run(), ctx, tb)
    return ctx.timer_end()


def roofline(w, ms_per_launch, peaks):
    if w["bound"] == "hbm":
        ach = w["bytes"] * w["batch"] / (ms_per_launch * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"]}
    ach = w["flops"] * w["batch"] / (ms_per_launch * 1e-3) / 1e12
    peak = peaks.get("fp64_tflops") or 37.0
    src = "measured cuBLAS DGEMM 8192^3 (profiles/r01_fp64_peak.json)" if peaks.get("fp64_tflops") else "nominal FP64 37 TFLOP/s (no measured DGEMM peak yet)"
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": src}


# ---- CPU arm: the oracle port on the host cores -------------------------------------------------
class CpuWork:
    """The oracle port on one bounded sample of the workload (inputs generated once, untimed)."""

    def __init__(self, w):
        import oracle
        self.oracle = oracle
        dt = {"f64": oracle.F64, "f32": oracle.F32, "i64": oracle.I64}[w["dtype"]]
        assert w["op"] == "matmul"
        self.m, self.k, self.n = w["m"], w["k"], w["n"]
        self.B = min(w["batch"], 1 << 24)
        self.a = oracle.fill(dt, oracle.SEED_A, 1, self.B * self.m * self.k)
        self.b = oracle.fill(dt, oracle.SEED_B, 1, self.B * self.k * self.n)
        self.threads = os.cpu_count() or 1
        self.one_pass()  # warm (page-in, thread start)

    def one_pass(self):
        self.oracle.matmul(self.a, self.b, self.m, self.k, self.n, threads=self.threads)

    def run_for(self, min_seconds, max_reps=200):
        reps, t0 = 0, time.perf_counter()
        while True:
            self.one_pass()
            reps += 1
            el = time.perf_counter() - t0
            if el >= min_seconds or reps >= max_reps:
                return self.B * reps / el, reps, el


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    cw = CpuWork(w)
    for _ in range(max(0, args.warmup)):
        cw.one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cw.one_pass()
    el = time.perf_counter() - t0
    value = cw.B * args.steps / el
    sample = f"{args.steps} steps x {cw.B} items of the workload, C++ oracle port, {cw.threads} threads"
    line = {"impl": "reference", "metric": "batched matrices/sec", "value": value, "unit": "matrices/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "note": "reference = CPU path; the Haskell library cannot be built here "
                       "(no ghc), so the C++ restatement (oracle port, unboxed: faster than boxed Data.Vector) is timed"},
            "cpu_baseline": {"value": value, "unit": "matrices/s", "cores": cw.threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---- our arm ---------------------------------------------------------------------------------------
def ours(args):
    import tensor_b200 as tb
    from tensor_b200 import distributed as D

    rank, local_rank, world = D.init_process_group()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = None
    if world > 1:
        import torch
        dev = torch.device("cuda", local_rank)
    ctx = tb.Context([local_rank])
    peaks = load_peaks()
    w = WORKLOADS[args.workload]
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    sampler = ClockSampler(local_rank)

    # ---- device-resident throughput (`value`) ------------------------------------------------
    ring = StepRing(ctx, tb, w)
    step = ring.first
    for _ in range(warmup):
        check_rc(ring.run(), ctx, tb)
    ctx.sync()
    D.barrier()
    n0 = ctx.launch_count()
    sampler.start()
    ctx.timer_begin()
    for _ in range(steps):
        check_rc(ring.run(), ctx, tb)
    ms = ctx.timer_end()
    launches = ctx.launch_count() - n0
    sampler.stop()
    D.barrier()
    ms_max = D.max_over_ranks(ms, dev)
    ms_per_step = ms_max / steps
    value = w["batch"] * world * steps / (ms_max * 1e-3)
    roof = roofline(w, ms / steps, peaks)
    roof["traffic"] = load_traffic(args.workload)
    clocks = sampler.summary()

    # ---- end to end through the host-buffer C-ABI entry (`e2e`) ----------------------------------
    e2e = None
    if w["op"] == "matmul" and not args.no_e2e:
        m, k, n = w["m"], w["k"], w["n"]
        dt = NPD[w["dtype"]]
        B = w["batch"]
        ha, hb, hc = tb.PinnedArray((B * m * k,), dt), tb.PinnedArray((B * k * n,), dt), tb.PinnedArray((B * m * n,), dt)
        # the host operands are the same synthetic stream (generated on the device, copied out once, untimed)
        L = tb._lib.lib()
        check_rc(L.t5_download(step.a._h, ha.array.ctypes.data_as(ctypes.c_void_p), ha.nbytes), ctx, tb)
        check_rc(L.t5_download(step.b._h, hb.array.ctypes.data_as(ctypes.c_void_p), hb.nbytes), ctx, tb)
        e_steps = max(2, min(steps, 8))
        for _ in range(2):
            tb.matmul_host(ctx, ha.array, hb.array, m, k, n, hc.array)
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            tb.matmul_host(ctx, ha.array, hb.array, m, k, n, hc.array)   # returns after the D2H of C completed
        el = time.perf_counter() - t0
        D.barrier()
        el = D.max_over_ranks(el, dev)
        # spot-check the host result against the device-resident result
        ref = step.c.to_numpy().reshape(-1)
        if not np.array_equal(ref[: 1 << 16].view(np.uint32 if dt == np.float32 else np.uint64),
                              hc.array[: 1 << 16].view(np.uint32 if dt == np.float32 else np.uint64)):
            raise SystemExit("host-entry result differs from device-entry result")
        e2e = {"value": B * world * e_steps / el, "unit": "matrices/s", "h2d_bytes_per_step": int(ha.nbytes + hb.nbytes),
               "d2h_bytes_per_step": int(hc.nbytes), "steps": e_steps, "api": "t5_matmul_host (pinned host buffers, 3-slot H2D/kernel/D2H pipeline)"}
        for p in (ha, hb, hc):
            p.free()
    ring.free()

    # ---- the other BASELINE configs (explanatory; not the headline) ----------------------------------
    others = {}
    if args.all and world == 1:
        for name in EXTRAS + (["matmul128_f64_64K"] if args.big else []):
            ww = WORKLOADS[name]
            try:
                st = StepRing(ctx, tb, ww)
                t = time_steps(ctx, tb, st, 30, 5) / 30
                r = roofline(ww, t, peaks)
                others[name] = {"matrices_per_s": ww["batch"] / (t * 1e-3), "ms": t, "achieved": r["achieved"],
                                "unit": r["unit"], "frac": r["frac"], "operand_sets": len(st.sets)}
                st.free()
            except Exception as e:  # report, never hide
                others[name] = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and w["op"] == "matmul":
        cw = CpuWork(w)
        v, reps, el = cw.run_for(10.0)
        cpu = {"value": v, "unit": "matrices/s", "cores": cw.threads, "kind": "port",
               "sample": f"{reps} passes over {cw.B} items of the same workload ({el:.1f} s), C++ oracle port of the reference semantics"}

    if rank == 0:
        line = {"metric": "batched matrices/sec", "value": value, "unit": "matrices/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
                "config": {"workload": w["desc"], "batch_per_gpu": w["batch"], "layout": "AoS (reference wire order)",
                           "l2": "%d operand set(s) x %.2f GB used round-robin: every launch streams from HBM (>> 126 MB L2), no flush"
                                 % (len(ring.sets), w["bytes"] * w["batch"] / 1e9),
                           "parallelism": f"batch-sharded x{world}, no collective on the data path"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if others:
            line["workloads"] = others
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--all", action="store_true", default=True, help="also time the other BASELINE configs (N=1)")
    ap.add_argument("--headline-only", dest="all", action="store_false")
    ap.add_argument("--big", action="store_true", help="include the 25.8 GB 128x128 Double config in --all")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
