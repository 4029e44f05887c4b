"""bench.py's contract, the parts that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, on all host
cores even when torchrun exports OMP_NUM_THREADS=1, with the same `config` as the GPU arm; the synthetic-input generator is
deterministic and independent of oracle/; the committed ncu traffic record belongs to the CUDA sources in the tree."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    return bench


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    bench = _bench()
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["n_gpus"] == 2 and line["steps"] == 1 and line["dtype"] == "f64"
    assert line["config"] == bench.config_of("HS")                      # identical in both arms, key by key
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "rows per step" in cb["sample"]
    from oracle.build import has_openmp
    assert cb["cores"] == (os.cpu_count() if has_openmp() else 1)       # not the single thread torchrun's environment asks for
    assert 0.0 < line["value"] < 100.0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3

    # the other ranks of a torchrun launch exit 0 without work and without output
    env["RANK"] = "1"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_synthetic_inputs_are_deterministic_and_follow_the_survey_recipe():
    bench = _bench()
    X, Y, phi0 = bench.synth(3000, 64, 64)
    X2, Y2, phi02 = bench.synth(3000, 64, 64)
    assert np.array_equal(Y, Y2) and np.array_equal(phi0, phi02) and np.array_equal(X, np.linspace(0, 10, 64)[:, None])
    np.random.seed(0)
    assert np.array_equal(phi0, np.random.randn(3000, 64))              # collapsed_mixture.py:41 with np.random.seed(0)
    from oracle import gpy_shim as G                                    # the generator's covariances are GPy's formulas
    assert np.array_equal(bench._rbf(X, 1.0, 2.0), G.RBF(1, 1.0, 2.0).K(X))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "oracle" not in src[src.index("def synth("):src.index("def measured_peaks(")]


def test_committed_traffic_record_belongs_to_the_sources_in_the_tree():
    """roofline.traffic is read from profiles/ncu_traffic_r02.json; the record carries the sha16 of the CUDA sources the captured library
    was built from.  A change under gpclust_b200/csrc or include/ needs a new `ncu --set full` capture (tools/ncu_traffic.py)."""
    bench = _bench()
    with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as f:
        rec = json.load(f)
    assert rec["src_sha16"] == bench.source_sha16(), "CUDA sources changed since the committed ncu capture: re-capture and re-run tools/ncu_traffic.py"
    val, src = bench.ncu_traffic("k_mohgp_loop", 1000000, 22, 20)
    assert "the same sources" in src
    alg = 20 * 8e6 * (64 + 4 * 64) + 22 * 8e6 * (64 + 5 * 64)
    assert 0.98 < val / alg < 1.03                                      # DRAM traffic = algorithmic bytes: nothing is re-read
