"""Static instruction counts per kernel from the built library (cuobjdump -sass): which pipes / mechanisms each kernel uses.
Usage: sass_mnemonics.py [lib.so] > profiles/<round>_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["DMMA", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "WARPSYNC", "CALL", "ACQBULK", "ATOM", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spamtree_b200", "libspamtree_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("st::", "").replace("void ", "")
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for c in COLS:
                if op.startswith(c):
                    cur[c] += 1
                    break
    print("cuobjdump -sass spamtree_b200/libspamtree_b200.so   (sm_100a; static instruction counts per kernel; tools/sass_mnemonics.py)")
    print("DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); LDGSTS = cp.async global->shared; DFMA/DMUL/DADD = FP64 FMA pipe; MUFU = rsqrt/rcp seeds;")
    print("BAR = CTA / named barriers; ACQBULK = griddepcontrol.wait (programmatic dependent launch); ATOM/RED = atomics.  tcgen05 / UTMA do not appear:")
    print("the path is FP64 (no tcgen05 FP64 kind) and its staged operands are ragged row blocks (DESIGN.md, 'Why not tcgen05/TMEM/TMA').\n")
    print(f"{'kernel':36s}" + "".join(f"{c:>9s}" for c in COLS) + f"{'total':>9s}")
    for k, c in per.items():
        print(f"{k[:36]:36s}" + "".join(f"{c[x]:9d}" for x in COLS) + f"{c['total']:9d}")


if __name__ == "__main__":
    main()
