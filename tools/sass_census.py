#!/usr/bin/env python
"""Static census of the shipped library's SASS (no GPU needed):  cuobjdump -sass libgpclust_b200.so | python tools/sass_census.py

Per kernel: instruction count and the counts of the mnemonics that identify the hardware paths this design rests on --
DMMA (FP64 tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier), STAS (st.async into distributed shared memory), UCGABAR
(cluster barrier), DFMA/DADD/DMUL (FP64 pipe), MUFU, LDG/STG/LDS/STS, STL/LDL (spills), BAR, RED/ATOM."""
import collections
import re
import sys

ABSENT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "HGMMA"]
KEYS = ["DMMA", "UBLKCP", "SYNCS", "STAS", "UCGABAR", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "RED", "ATOM", "SHFL"]


def main():
    fn, per, arch = None, collections.OrderedDict(), set()
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
    for line in sys.stdin:
        if "arch =" in line:
            arch.add(line.split("=")[1].strip())
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = ins.match(line)
        if m and fn:
            op = m.group(1)
            per[fn]["_all"] += 1
            hit = [k for k in KEYS + ABSENT if op.startswith(k)]
            per[fn][hit[0] if hit else op] += 1
    print("arch: %s; %d kernels" % (", ".join(sorted(arch)), len(per)))
    print("%-64s %8s " % ("kernel (demangled name shortened)", "instr") + " ".join("%6s" % k for k in KEYS))
    tot = collections.Counter()
    for fn, c in per.items():
        name = re.sub(r"^_Z\w*?\d+gpc\d+", "", fn)[:64]
        print("%-64s %8d " % (name, c["_all"]) + " ".join("%6d" % c[k] for k in KEYS))
        tot.update(c)
    print("%-64s %8d " % ("TOTAL", tot["_all"]) + " ".join("%6d" % tot[k] for k in KEYS))
    absent = [k for k in ABSENT if tot[k] == 0]
    print("absent (tcgen05 / tensor-map TMA / wgmma have no FP64 form; see DESIGN.md section 4): " + ", ".join(absent))


if __name__ == "__main__":
    main()
