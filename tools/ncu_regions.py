"""Split the stall samples of an `ncu --page source --csv --print-source sass` export by execution-count class and opcode.
Usage: python tools/ncu_regions.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        s = int(r[idx["Warp Stall Sampling (All Samples)"]])
        n = int(r[idx["Instructions Executed"]])
    except ValueError:
        continue
    data.append((s, n, r[idx["Source"]].strip(), {h: int(r[idx[h]] or 0) for h in reasons}))
T = sum(d[0] for d in data)
NI = sum(d[1] for d in data)
print("samples", T, "warp instructions", NI)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for s, n, src, st in data:
    a = agg[n]
    a[0] += s
    a[1] += 1
    for k, v in st.items():
        a[2][k] += v
for n, (s, c, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:6]:
    print("exec count %9d : %4d instrs, %5.1f%% of samples | %s" % (n, c, 100.0 * s / T, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / T) for k, v in st.most_common(5))))
ops = collections.Counter()
ins = collections.Counter()
for s, n, src, st in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else src[:12]).split(".")[0]
    ops[op] += s
    ins[op] += n
for op, s in ops.most_common(14):
    print("  %-10s samples %5.1f%%  instr %5.1f%% (%d)" % (op, 100.0 * s / T, 100.0 * ins[op] / NI, ins[op]))
if len(sys.argv) > 2:
    for s, n, src, st in sorted(data, key=lambda d: -d[0])[:int(sys.argv[2])]:
        print("  %5.2f%% n=%9d %s | %s" % (100.0 * s / T, n, src[:70], ", ".join("%s %d" % (k[6:], v) for k, v in collections.Counter(st).most_common(3))))
