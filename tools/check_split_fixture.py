#!/usr/bin/env python
"""Replay tests/golden/reference_splits_*.npz -- the reference's own MOHGP merge-split search (collapsed_mixture.py:96-205), recorded by
tests/golden/make_reference_goldens.py -- through the GPU path and print one JSON record per fixture: steepest trace, the clusters
attempted, accept / reject of every split with its bound increment, final K, bound, assignments and the state of the legacy numpy
stream afterwards.  Needs no oracle (the fixture is the reference's output).

    python tools/check_split_fixture.py [fixture.npz ...]        exit code 0 when every record says ok
"""
import contextlib
import glob
import io
import json
import os
import re
import sys
import time

T0 = time.perf_counter()
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def kern_of(spec, mod):
    ctor = {"rbf": mod.RBF, "matern32": mod.Matern32, "matern52": mod.Matern52}
    parts = [ctor[k](1, v, l) if k in ctor else (mod.Bias(1, v) if k == "bias" else mod.White(1, v)) for k, v, l in spec]
    out = parts[0]
    for p in parts[1:]:
        out = out + p
    return out.fix()


def replay(path):
    from gpclust_b200 import MOHGP, kern
    d = np.load(path)
    seed = {"reference_splits_mohgp.npz": 7, "reference_splits_mohgp_dp.npz": 8}[os.path.basename(path)]
    m = MOHGP(d["X"], kern_of(json.loads(str(d["kernF"])), kern), kern_of(json.loads(str(d["kernY"])), kern), d["Y"], K=int(d["K"]),
              alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]), phi_init=d["phi0"])
    rec = {"fixture": os.path.basename(path)}
    rec["bound0_rel"] = abs(m.bound() - float(d["bound0"])) / abs(float(d["bound0"]))
    m.optimize(method="steepest", maxiter=40, verbose=False)
    tr, ref = np.array(m.optimize_trace, dtype=np.float64).reshape(-1, 4), d["trace_steepest"]
    rec["trace_iterations_equal"] = bool(tr.shape == ref.shape and np.array_equal(tr[:, 0], ref[:, 0]))
    rec["trace_bound_rel"] = float(np.max(np.abs(tr[:, 1] - ref[:, 1]) / np.abs(ref[:, 1]))) if rec["trace_iterations_equal"] else None
    rec["bound1_rel"] = abs(m.bound() - float(d["bound1"])) / abs(float(d["bound1"]))
    np.random.seed(seed + 2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.systematic_splits(verbose=True)
    events = re.findall(r"attempting to split cluster\s+(\d+)|split (failed|suceeded), bound changed by:\s+(\S+)", buf.getvalue().replace("\r", "\n"))
    attempts = [int(a) for a, _, _ in events if a]
    results = np.array([(1.0 if r == "suceeded" else 0.0, float(v)) for a, r, v in events if r], dtype=np.float64).reshape(-1, 2)
    rec["attempts"], rec["attempts_equal"] = attempts, attempts == [int(v) for v in d["split_attempts"]]
    rec["decisions_equal"] = bool(results.shape == d["split_results"].shape and np.array_equal(results[:, 0], d["split_results"][:, 0]))
    rec["increment_max_abs_diff"] = float(np.max(np.abs(results[:, 1] - d["split_results"][:, 1]))) if rec["decisions_equal"] else None
    rec["K"], rec["K_equal"] = int(m.K), int(m.K) == int(d["K2"])
    rec["bound2_rel"] = abs(m.bound() - float(d["bound2"])) / abs(float(d["bound2"]))
    rec["argmax_equal"] = bool(rec["K_equal"] and np.array_equal(np.argmax(m.phi, 1), np.argmax(d["phi2"], 1)))
    rec["rng_state_equal"] = bool(np.array_equal(np.random.randn(3), d["rng_after"]))
    rec["ok"] = bool(rec["bound0_rel"] < 1e-9 and rec["trace_iterations_equal"] and rec["trace_bound_rel"] < 1e-9 and rec["bound1_rel"] < 1e-9 and
                     rec["attempts_equal"] and rec["decisions_equal"] and rec["increment_max_abs_diff"] < 1e-5 and rec["K_equal"] and
                     rec["bound2_rel"] < 1e-8 and rec["argmax_equal"] and rec["rng_state_equal"])
    return rec


def main():
    paths = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "reference_splits_*.npz")))
    ok = True
    for p in paths:
        t = time.perf_counter()
        rec = replay(p)
        rec["seconds"], rec["seconds_since_start"] = time.perf_counter() - t, time.perf_counter() - T0
        print(json.dumps(rec), flush=True)
        ok = ok and rec["ok"]
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
