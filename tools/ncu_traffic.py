"""profiles/ncu_traffic_r02.json from an `ncu --page raw --csv` export of the shipped binary: DRAM bytes (read + write) per G phase and per
S phase of the dominant kernel, and the sha16 of the CUDA sources the captured library was built from (bench.py reports `roofline.traffic` from it and
says whether the running library is built from the same sources).
Usage: python tools/ncu_traffic.py raw.csv ROWS EVALS G_PHASES source-name"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, rows, evals, g_phases, source = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
table = list(csv.reader(open(raw)))
hdr, units = table[0], table[1]
out = {"rows": rows, "source": source, "kernels": {}}
sys.path.insert(0, ROOT)
from bench import source_sha16  # noqa: E402  (hash of the CUDA sources + the C-ABI header: what the library is built from)
out["src_sha16"] = source_sha16()


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


for r in table[2:]:
    name = r[hdr.index("Kernel Name")]
    i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    total = to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w])
    dur = r[hdr.index("gpu__time_duration.sum")]
    if "k_mohgp_loop" in name:
        # one launch = g_phases G phases + evals S phases; algorithmic split 8n(D+4K) : 8n(D+5K) at D = K
        g_alg, s_alg = 5.0, 6.0
        per_unit = total / (g_phases * g_alg + evals * s_alg)
        out["kernels"]["k_mohgp_loop"] = {"dram_bytes_per_g_phase": per_unit * g_alg, "dram_bytes_per_s_phase": per_unit * s_alg,
                                          "dram_bytes_launch": total, "evaluations": evals, "g_phases": g_phases, "duration": dur + " " + units[hdr.index("gpu__time_duration.sum")]}
    elif "k_natgrad" in name:
        out["kernels"]["k_natgrad"] = {"dram_bytes_per_g_phase": total, "dram_bytes_per_s_phase": 0.0, "duration": dur}
    elif "k_step_stats" in name:
        out["kernels"]["k_step_stats"] = {"dram_bytes_per_g_phase": 0.0, "dram_bytes_per_s_phase": total, "duration": dur}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
