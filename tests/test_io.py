"""Ingest helpers: loaders for the reference's data files and the notebook's preparation steps."""
import ctypes
import os

import numpy as np
import pytest

from helpers import GOLDEN


def test_loaders_read_the_reference_files():
    from gpclust_b200 import io
    X, Y, z = io.load_split_data(os.path.join(GOLDEN, "split_data_test.csv"))
    assert X.shape == Y.shape == (z.shape[0], 1) and set(np.unique(z)) <= {0, 1} and X.shape[0] > 100
    sp, rep, t = io.load_kalinka_pdata(os.path.join(GOLDEN, "kalinka09_pdata.csv"))
    assert len(sp) == len(rep) == len(t) and len(t) > 50
    sp1, rep1, t1 = io.load_kalinka_pdata(os.path.join(GOLDEN, "kalinka09_pdata.csv"), species=sp[0])
    assert len(t1) < len(t) and np.all(np.diff(np.sort(np.unique(t1))) > 0)
    a = io.load_numpy_fixture(os.path.join(GOLDEN, "twoD_clustering_example.npy"))
    assert a.shape == (506, 2)


@pytest.mark.gpu
def test_normalisation_and_replicate_scores_match_numpy():
    from gpclust_b200 import io
    rng = np.random.RandomState(0)
    times = np.repeat(np.arange(2.0, 16.0, 2.0), 8)  # 7 time points x 8 replicates
    Y = rng.randn(3000, times.size) * rng.rand(3000, 1) * 3 + rng.randn(3000, 1)
    e = Y - Y.mean(1)[:, None]
    e /= e.std(1)[:, None]
    got = io.normalise_rows(Y)
    assert np.max(np.abs(got - e)) < 1e-13
    means, stds = [], []
    for t in np.unique(times):
        idx = times == t
        means.append(e[:, idx].mean(1))
        stds.append(e[:, idx].std(1))
    means, stds = np.vstack(means).T, np.vstack(stds).T
    scores = means.std(1) / stds.mean(1)
    got = io.replicate_scores(e, times)
    assert np.max(np.abs(got - scores) / scores) < 1e-12
    top, index = io.select_top_series(e, times, n=500)
    assert np.array_equal(index, np.argsort(scores)[::-1][:500]) and top.shape == (500, times.size)


@pytest.mark.gpu
def test_pipelined_ingest_many_chunks_back_to_back(monkeypatch):
    """Host -> device ingest (_cabi._Ingest): chunked staging / DMA / layout conversion.  Two uploads back to back through the same
    staging ring with many small chunks each (the second must not overwrite staging buffers the first one's DMA is still reading),
    ragged last chunk, both layouts; the device content must equal the host arrays exactly."""
    import torch
    from gpclust_b200 import _cabi
    from gpclust_b200._cabi import lib, check, ptr
    ing = _cabi._Ingest()
    monkeypatch.setattr(ing, "CHUNK_BYTES", 1 << 16)  # 64 KB chunks: 128 rows of 64 doubles
    dev = torch.device("cuda", torch.cuda.current_device())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.RandomState(0)
    for rep in range(3):
        outs = []
        for (N, C, CP, layout) in [(5000 + 37 * rep, 64, 64, "features"), (5000 + 37 * rep, 64, 64, "logits"), (3001, 24, 24, "features"), (3001, 7, 8, "logits")]:
            src = rng.randn(N, C)
            if rep == 2:
                src = torch.from_numpy(src).pin_memory().numpy()  # page-locked input: DMA straight out of it (no staging copy)
            dst = torch.full((_cabi.padded_rows(N), CP), 7.0, dtype=torch.float64, device=dev)
            ing.upload_pack(src, C, CP, _cabi.LAYOUT[layout], dst, dev, st)
            outs.append((src, dst, N, C, CP, layout))
        for src, dst, N, C, CP, layout in outs:  # verified only after ALL uploads were enqueued
            back = torch.empty((N, C), dtype=torch.float64, device=dev)
            check(lib.gpc_unpack_rows(ptr(dst), N, C, CP, _cabi.LAYOUT[layout], ptr(back), st), "gpc_unpack_rows")
            assert np.array_equal(back.cpu().numpy(), src), (N, C, layout, rep)
