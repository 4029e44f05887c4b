"""Ingest helpers: the step before the path.

The reference has no loader module; its notebooks read the inputs with numpy and prepare them inline
(`notebooks/drosophila.ipynb` cells [2] and [4], `notebooks/OMGP_demo.ipynb` cell [2], `mixture of Gaussians.ipynb` cell [2]).
These functions restate those cells; the two N-sized steps (row normalisation, replicate scores) run on the device so that
a 16M-series matrix never makes an extra round trip through host temporaries.
"""
import ctypes
import pickle

import numpy as np

from . import _cabi
from ._cabi import lib, check, ptr


def load_numpy_fixture(path):
    """`notebooks/twoD_clustering_example.numpy`: a Python-2 pickle of a float64 array (`np.load` in the notebook)."""
    with open(path, "rb") as f:
        head = f.read(6)
        f.seek(0)
        if head.startswith(b"\x93NUMPY"):
            return np.load(f)
        return np.asarray(pickle.load(f, encoding="latin1"), dtype=np.float64)


def load_split_data(path):
    """`data/split_data_test.csv` (OMGP_demo.ipynb cell [2]) -> (X (N,1), Y (N,1), true_assignment (N,))."""
    a = np.loadtxt(path, delimiter=",", skiprows=1)
    return a[:, 1, None].copy(), a[:, 2, None].copy(), a[:, 0].astype(int)


def load_kalinka_pdata(path, species=None):
    """`data/kalinka09_pdata.csv`: rows (species code, replicate, hours) of the expression matrix's columns.
    -> (species (D,), replicates (D,), times (D,)); `species='me'` keeps one species' columns (drosophila.ipynb uses mel)."""
    sp, rep, t = [], [], []
    with open(path) as f:
        for line in f:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 3 or not parts[1].replace(".", "").isdigit():
                continue
            sp.append(parts[0])
            rep.append(float(parts[1]))
            t.append(float(parts[2]))
    sp, rep, t = np.array(sp), np.array(rep), np.array(t)
    if species is not None:
        keep = sp == species
        sp, rep, t = sp[keep], rep[keep], t[keep]
    return sp, rep, t


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(a, torch):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(_cabi.f64(a)).cuda()


def normalise_rows(Y, as_tensor=False):
    """expression -= mean(1); expression /= std(1)  (drosophila.ipynb cell [2]) on the device."""
    torch = _cabi.require_cuda()
    y = _dev(Y, torch)
    out = torch.empty_like(y)
    check(lib.gpc_normalise_rows(ptr(y), y.shape[0], y.shape[1], ptr(out), _stream(torch)), "gpc_normalise_rows")
    return out if as_tensor else _cabi.to_host(out)


def replicate_scores(Y, times):
    """means.std(1) / stds.mean(1) over the replicate groups of equal `times` (drosophila.ipynb cell [4])."""
    torch = _cabi.require_cuda()
    y = _dev(Y, torch)
    uniq, group = np.unique(np.asarray(times), return_inverse=True)
    if len(uniq) > 64:
        raise ValueError("at most 64 distinct time points")
    g = torch.from_numpy(group.astype(np.int32)).cuda()
    out = torch.empty(y.shape[0], dtype=torch.float64, device="cuda")
    check(lib.gpc_replicate_scores(ptr(y), y.shape[0], y.shape[1], ptr(g), len(uniq), ptr(out), _stream(torch)), "gpc_replicate_scores")
    return _cabi.to_host(out)


def select_top_series(Y, times, n=500):
    """Rank the series by replicate score and keep the top n (drosophila.ipynb cell [4]) -> (Y[index], index)."""
    index = np.argsort(replicate_scores(Y, times))[::-1][:n]
    return np.asarray(Y)[index], index
