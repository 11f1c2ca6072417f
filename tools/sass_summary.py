#!/usr/bin/env python3
"""profiles/sass_summary.txt: static SASS opcode counts of the built library (cuobjdump -sass), totals and the hot kernels.
usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "shll_sve_cfd_b200", "libshll_b200.so")
OPS = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "DMUL", "MUFU", "LOP3", "FMNMX", "SHFL",
       "STG", "STS", "LDS", "MEMBAR"]
HOT = [r"step2d_acc2_kernelILi0ELi12E", r"step1d_acc2_kernelILi1ELi0ELi4E", r"step2d_acc_kernelILi1ELi0ELi0ELi16ELi0E", r"step2d_acc_kernelILi2ELi1ELi0ELi12ELi0E", r"step2d_acc_kernelILi2ELi1ELi1ELi12ELi0E",
       r"step1d_acc_kernelILi1ELi0ELi6E", r"step1d_kernelILi1ELi0ELi0ELi1ELi2ELb1E", r"step1d_kernelILi2ELi1ELi0ELi0ELi2ELb1E",
       r"step1d_kernelILi1ELi0ELi0ELi0ELi1ELb1E", r"persist1d_kernelILi2ELi1ELi0ELi1ELi2ELb1E", r"persist1d_kernelILi2ELi1ELi0ELi0ELi2ELb1E",
       r"step2d_tma_kernelILi1ELi0ELi0ELi0ELi1ELb1E", r"step2d_tma_kernelILi2ELi1ELi0ELi0ELi1ELb1E"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kern, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            counts[kern][m.group(1).split(".")[0]] += 1
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# SASS evidence: static opcode counts of shll_sve_cfd_b200/libshll_b200.so (cuobjdump -sass); made by tools/sass_summary.py")
    print("# UTMALDG / UTMASTG = TMA load / store (cp.async.bulk.tensor), SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 = packed FP32x2,")
    print("# DFMA / DMUL = the two FP64 islands of STRICT mode.  No tensor-core opcodes: the path has no contraction.")
    print(f"architectures in the fat binary: {', '.join(archs)}")
    print(f"TOTAL over {len(counts)} kernels: " + "  ".join(f"{o}={tot[o]}" for o in OPS if tot[o]))
    print()
    for pat in HOT:
        for k in sorted(counts):
            if re.search(pat, k):
                name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
                c = counts[k]
                print(name[:120])
                print(f"    instructions={sum(c.values())}  " + "  ".join(f"{o}={c[o]}" for o in OPS if c[o]))


if __name__ == "__main__":
    main()
