"""Machine-code evidence for the hardware claims in DESIGN.md: disassembles tensor_b200/libt5b200.so (cuobjdump
-sass) and counts, per kernel, the mnemonics that prove which units a kernel uses:
  UTCHMMA / UTCBAR / LDTM / UTCATOMSWS     tcgen05.mma, tcgen05.commit, tcgen05.ld, TMEM allocation (5th-gen tensor cores)
  UTMALDG / UTMASTG / SYNCS                TMA tensor copies and their mbarriers
  DMMA                                     FP64 tensor-core mma.sync.m8n8k4
  LDGSTS                                   cp.async global->shared
  DFMA / DMUL / DADD, FFMA / FMUL / FADD   fused vs unfused arithmetic (the exact kernels must not fuse)
  FMUL2 / FADD2 / FFMA2                    packed FP32 pairs
Writes profiles/sass_<family>.txt for the kernel families named on the command line (substring match on the demangled
kernel name) plus profiles/sass_summary.txt.  Usage: python tools/sass_evidence.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tensor_b200", "libt5b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "HMMA", "LDGSTS",
         "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "FMUL2", "FADD2", "FFMA2", "IMAD", "SHFL", "BAR", "LDS", "STS",
         "LDG", "STG", "LDL", "STL", "MUFU", "ATOMG", "ATOMS", "REDG"]
FAMILIES = {
    "headline_tile_matmul4x4_f32": "tile_kernel<t5::MatmulOp<float, 4, 4, 4>",
    "dmma_tma": "dmma_tma_kernel",
    "tf32x3_tcgen05": "tf32x3_kernel",
    "xgemm_exact": "xgemm_kernel",
    "lu_cta_per_matrix": "lu_kernel",
    "inverse8_f64": "InverseOp<double, 8>",
    "det9_reg_f64": "DetOp<double, 9>",
    "longdot_rows": "longdot_rows_kernel",
    "relayout": "relayout_",
    "soa_prod_split": "soa_prod_split",
    "staged_gather": "staged_gather_kernel",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per_kernel = collections.OrderedDict()
    cur, it = None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it)
            per_kernel[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            per_kernel[cur][op] += 1
            if op in ("DMMA", "UTCHMMA", "LDTM", "UTMALDG", "LDGSTS"):
                per_kernel[cur][op + m.group(2)] += 1
            per_kernel[cur]["_total"] += 1
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    total = collections.Counter()
    for c in per_kernel.values():
        total.update(c)
    with open(os.path.join(ROOT, "profiles", f"sass_summary_{tag}.txt"), "w") as f:
        f.write(f"libt5b200.so built from {sha} (+ working tree), {len(per_kernel)} kernels, sm_100a; counts over the whole library\n")
        for op in WATCH:
            f.write(f"{op:12s} {total[op]}\n")
        f.write("\nkernels that contain the tensor-core / TMA mnemonics:\n")
        for k, c in per_kernel.items():
            hits = {op: c[op] for op in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "DMMA", "FMUL2", "FADD2", "FFMA2") if c[op]}
            if hits:
                f.write(f"  {k[:150]}\n      {hits}\n")
    for fam, pat in FAMILIES.items():
        ks = [(k, c) for k, c in per_kernel.items() if pat in k]
        if not ks:
            continue
        with open(os.path.join(ROOT, "profiles", f"sass_{fam}_{tag}.txt"), "w") as f:
            f.write(f"cuobjdump -sass tensor_b200/libt5b200.so | kernels matching '{pat}' ({len(ks)}); built from {sha} (+ working tree)\n")
            for k, c in ks:
                f.write(f"\n{k}\n  instructions: {c['_total']}\n")
                for op, v in sorted(c.items(), key=lambda kv: -kv[1]):
                    if op != "_total" and (op.split(".")[0] in WATCH):
                        f.write(f"  {op:40s} {v}\n")
    print("wrote profiles/sass_*_%s.txt for %d kernels" % (tag, len(per_kernel)))


if __name__ == "__main__":
    main()
