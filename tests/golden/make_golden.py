"""Extract the reference's only seeded known-answer outputs into small committed fixtures.

Run in the build container (where /root/reference is mounted); the GPU box never reads
/root/reference. Sources:
  notebooks/mixture of Gaussians.ipynb  (stored stdout of cells [8], [11], [13], [21], [23])
  notebooks/twoD_clustering_example.numpy  (py2 pickle of a (506, 2) float64 array)
  data/split_data_test.csv, data/kalinka09_pdata.csv  (demo inputs without goldens)
Outputs (all under tests/golden/):
  mog_notebook_golden.json, twoD_clustering_example.npy, split_data_test.csv, kalinka09_pdata.csv
"""
import json
import os
import pickle
import re
import shutil

import numpy as np

REF = os.environ.get("GPCLUST_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

LINE = re.compile(r"iteration (\d+) bound=(\S+) grad=(\S+), beta=(\S+)")


def cell_text(cell):
    out = []
    for o in cell.get("outputs", []):
        if o.get("output_type") in ("stream", "pyout"):
            out.append("".join(o.get("text", "")))
    return "".join(out).replace("\r", "\n")


def parse_trace(text):
    """-> list of dict(iteration, bound, grad, beta) with the printed strings kept verbatim."""
    rows = []
    for m in LINE.finditer(text):
        rows.append({"iteration": int(m.group(1)), "bound": m.group(2), "grad": m.group(3), "beta": m.group(4)})
    return rows


def main():
    nb = json.load(open(os.path.join(REF, "notebooks", "mixture of Gaussians.ipynb")))
    cells = nb["worksheets"][0]["cells"]
    gold = {"source": "notebooks/mixture of Gaussians.ipynb", "cells": {}}
    for idx in (8, 11, 13, 21, 23):
        c = cells[idx]
        text = cell_text(c)
        gold["cells"][str(idx)] = {
            "prompt_number": c.get("prompt_number"),
            "input": "".join(c.get("input", "")),
            "stdout": text,
            "trace": parse_trace(text),
            "messages": [l.strip() for l in text.split("\n")
                         if l.strip() and not l.strip().startswith("iteration")],
        }
    # markers appended to an iteration line ("... beta=0.23 vb converged (ftol)") are kept in stdout
    with open(os.path.join(HERE, "mog_notebook_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)

    with open(os.path.join(REF, "notebooks", "twoD_clustering_example.numpy"), "rb") as f:
        X = pickle.load(f, encoding="latin1")
    X = np.ascontiguousarray(X, dtype=np.float64)
    assert X.shape == (506, 2)
    np.save(os.path.join(HERE, "twoD_clustering_example.npy"), X)
    for name in ("split_data_test.csv", "kalinka09_pdata.csv"):
        shutil.copyfile(os.path.join(REF, "data", name), os.path.join(HERE, name))
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
