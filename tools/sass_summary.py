#!/usr/bin/env python3
"""Per-kernel SASS opcode counts of libsympgpr_b200.so (cuobjdump -sass): the evidence that the shipped cubins are sm_100a
and which hardware paths each kernel uses -- DMMA (FP64 tensor core; tcgen05 has no f64 kind), UBLKCP (cp.async.bulk, the
TMA engine's 1-D path), LDGSTS (cp.async), SYNCS (mbarrier), DFMA/DMUL/DADD (FP64 pipe), MUFU, BAR, REDUX/SHFL.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sympgpr_b200", "libsympgpr_b200.so")
WATCH = ["DMMA", "UBLKCP", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL",
         "ATOM", "RED", "UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "IMMA", "HMMA"]


def strip_args(sig):
    """`void ns::k<(Layout)0, 1>(Args, long)` -> `ns::k<(Layout)0, 1>` (drop the return type and the parameter list)."""
    sig = sig.strip()
    if sig.endswith(")"):
        depth = 0
        for i in range(len(sig) - 1, -1, -1):
            if sig[i] == ")":
                depth += 1
            elif sig[i] == "(":
                depth -= 1
                if depth == 0:
                    sig = sig[:i]
                    break
    return sig[5:] if sig.startswith("void ") else sig


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w):
                    kernels[cur][w] += 1
                    break
    demangled = {}
    try:
        dm = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(kernels, dm))
    except Exception:
        pass
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)}  (cuobjdump -sass; architectures in the fat binary: {', '.join(archs)})")
    print("# counts are static instruction counts per kernel; columns with all zeros are omitted\n")
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    print("kernel".ljust(64), "instr".rjust(7), *[c.rjust(7) for c in cols])
    tot = collections.Counter()
    for name, cnt in kernels.items():
        nm = strip_args(demangled.get(name, name))
        nm = nm.replace("sgp::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(Layout)", "")
        print(nm[:64].ljust(64), str(cnt["_total"]).rjust(7), *[str(cnt[c]).rjust(7) for c in cols])
        tot.update(cnt)
    print("TOTAL".ljust(64), str(tot["_total"]).rjust(7), *[str(tot[c]).rjust(7) for c in cols])
    absent = [w for w in ("UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG") if not tot[w]]
    if absent:
        print(f"\n# not present: {', '.join(absent)} -- tcgen05.mma has no f64 kind (SURVEY 7), the FP64 tensor path of sm_100a is DMMA;")
        print("# bulk copies are the 1-D cp.async.bulk (UBLKCP), not tensor-map TMA (UTMALDG): operand slabs are 1 KB contiguous rows.")


if __name__ == "__main__":
    sys.exit(main())
