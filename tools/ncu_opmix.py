"""Aggregate an `ncu --page source --csv` export by SASS opcode: instruction share and stall-sample share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
ops = collections.Counter()
stall = collections.Counter()
tot = totst = 0
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    src = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    try:
        n = int(r[idx["Instructions Executed"]] or 0)
        s = int(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    except ValueError:
        continue
    ops[op] += n
    stall[op] += s
    tot += n
    totst += s
print("total warp instructions", tot, "stall samples", totst)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-12s %12d %5.1f%%  stall %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * stall[op] / max(totst, 1)))
