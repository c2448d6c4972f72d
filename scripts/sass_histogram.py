"""SASS opcode histogram of dcfp_b200/lib/libdcfp_b200.so per kernel family (for profiles/): shows that the shipped
kernels are sm_100a code that moves tiles with TMA (UTMALDG), waits on mbarrier transaction barriers (SYNCS), computes
with packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) and leaves CTAs through vector / fp64 reductions (REDG).  Needs no
GPU: `python scripts/sass_histogram.py > profiles/rNN_sass_opcodes.txt` in the build container."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dcfp_b200", "lib", "libdcfp_b200.so")
# opcode prefixes worth listing (everything else -- address arithmetic, moves, branches -- is summed as "other")
KEEP = ("UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "REDG", "RED.", "ATOMG", "ATOMS", "LDS.", "STS.", "LDG.", "STG.",
        "SHFL", "VOTE", "BAR.", "MEMBAR", "DADD", "DFMA", "DMUL", "F2FP", "PRMT", "HFMA2", "MUFU", "ERRBAR", "CCTL", "LDGSTS", "FFMA", "FADD")
FAMILY = re.compile(r"(\w+_kernel)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    fam, hist, size, variants = None, collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for line in out.splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            name = subprocess.run(["c++filt", s.split(":", 1)[1].strip()], capture_output=True, text=True).stdout.strip()
            m = FAMILY.search(name)
            fam = m.group(1) if m else name[:60]
            variants[fam] += 1
            continue
        m = re.match(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", s)
        if m and fam:
            op = m.group(1)
            size[fam] += 1
            hist[fam][op if op.startswith(KEEP) else "(other)"] += 1
    print("# SASS opcode histogram of dcfp_b200/lib/libdcfp_b200.so (cuobjdump -sass; every kernel is in an sm_100a cubin -- the one")
    print("# sm_52 ELF `cuobjdump -lelf` lists is nvcc's empty device-link stub), per kernel family, summed over")
    print("# the template instantiations: TMA tensor-tile loads (UTMALDG), mbarrier transaction barriers (SYNCS), packed")
    print("# fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2), reductions to the arenas (REDG: .F32x4 vector form from the fused BN backward,")
    print("# .F64 from the hook path).  Regenerate: python scripts/sass_histogram.py")
    for f in sorted(hist, key=lambda k: -size[k]):
        print("== %s  (%d instantiations, %d instructions)" % (f, variants[f], size[f]))
        for op, n in hist[f].most_common():
            if op != "(other)":
                print("%7d %s" % (n, op))
        print("%7d (other)" % hist[f]["(other)"])


if __name__ == "__main__":
    sys.exit(main())
