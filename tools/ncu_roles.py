"""Split the stall samples of an `ncu --page source --csv --print-source sass` export of ONE kernel into its LOOPS: maximal runs of
consecutive SASS instructions that were executed (about) the same number of times.  For every loop that took at least 1 % of the
samples: instructions, executions, share of the samples, the leading stall reasons and the opcode classes -- enough to tell the
warp roles of a warp-specialised kernel apart (tensor loops carry DMMA, element loops DFMA / SHFL / STG, producers UBLKCP / SYNCS).
Usage: python tools/ncu_roles.py src.csv [min_share_percent]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx["Warp Stall Sampling (All Samples)"]])
        n = int(r[idx["Instructions Executed"]])
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]].strip())
    op = m.group(2).split(".")[0] if m else "?"
    data.append((s, n, op, {h.replace("stall_", ""): int(r[idx[h]] or 0) for h in reasons}))
total = sum(d[0] for d in data)


def same(a, b):  # executed about equally often (predicated instructions count the same; tails differ by a few)
    return a > 0 and b > 0 and abs(a - b) <= 0.02 * max(a, b)


loops, cur = [], None
for i, (s, n, op, st) in enumerate(data):
    if cur is not None and same(cur["n"], n):
        cur["end"] = i
    else:
        cur = {"start": i, "end": i, "n": n}
        loops.append(cur)
# merge neighbouring runs separated by a few instructions of another count (branches, rarely taken paths)
merged = []
for L in loops:
    if merged and same(merged[-1]["n"], L["n"]) and L["start"] - merged[-1]["end"] <= 8:
        merged[-1]["end"] = L["end"]
    elif len(merged) >= 2 and same(merged[-2]["n"], L["n"]) and merged[-1]["end"] - merged[-1]["start"] < 8:
        merged[-2]["end"] = L["end"]
        merged.pop()
    else:
        merged.append(dict(L))
print("total stall samples %d, SASS instructions %d" % (total, len(data)))
print("%-7s %-7s %-10s %-7s  %-52s %s" % ("instrs", "first", "executed", "samples", "leading stall reasons (share of ALL samples)", "opcode classes (instructions in the loop)"))
for L in merged:
    seg = data[L["start"]:L["end"] + 1]
    s = sum(d[0] for d in seg)
    if s < total * min_share / 100.0:
        continue
    st = collections.Counter()
    ops = collections.Counter()
    for d in seg:
        for k, v in d[3].items():
            st[k] += v
        ops[d[2]] += 1
    cls = {"DMMA": ops["DMMA"], "FP64": ops["DFMA"] + ops["DADD"] + ops["DMUL"] + ops["DSETP"], "LDS": ops["LDS"], "STS": ops["STS"], "LDG": ops["LDG"],
           "STG": ops["STG"], "SHFL": ops["SHFL"], "SYNCS": ops["SYNCS"], "UBLKCP": ops["UBLKCP"], "BAR": ops["BAR"], "MUFU": ops["MUFU"]}
    print("%-7d %-7d %-10d %5.1f%%   %-52s %s" % (len(seg), L["start"], L["n"], 100.0 * s / total,
                                                  " ".join("%s %.1f" % (k, 100.0 * v / total) for k, v in st.most_common(4)),
                                                  " ".join("%s %d" % (k, v) for k, v in cls.items() if v)))
