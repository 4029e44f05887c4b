"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/gpclust_b200.h declares, the ctypes prototypes mirror the header, and compute entry points fail
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpclust_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|long long)\s+(gpc_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_header_symbols_exported_and_prototypes_match():
    from gpclust_b200 import _cabi
    decls = _declared()
    assert len(decls) >= 20
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name, nargs in decls.items():
        assert hasattr(lib, name), "symbol %s declared in the header is not exported" % name
        assert name in _cabi.PROTOTYPES, "no ctypes prototype for %s" % name
        assert len(_cabi.PROTOTYPES[name]) == nargs, (name, len(_cabi.PROTOTYPES[name]), nargs)
    assert set(_cabi.PROTOTYPES) == set(decls)


def test_geometry_queries_without_gpu():
    from gpclust_b200 import _cabi
    lib = _cabi.lib
    assert lib.gpc_abi_version() >= 1
    assert [lib.gpc_padded_k(k) for k in (1, 8, 9, 16, 17, 33, 64)] == [8, 8, 16, 16, 32, 64, 64]
    assert lib.gpc_padded_k(0) == -1 and lib.gpc_padded_k(513) == -1
    assert [lib.gpc_padded_k(k) for k in (65, 128, 129, 500, 512)] == [128, 128, 192, 512, 512]  # chunks of 64
    assert [lib.gpc_padded_d(d) for d in (1, 8, 56, 64)] == [8, 8, 56, 64] and lib.gpc_padded_d(65) == -1
    assert lib.gpc_padded_rows(1) == 64 and lib.gpc_padded_rows(1000000) == 1000000 and lib.gpc_padded_rows(65) == 128
    assert lib.gpc_stats_len(64, 64) == 64 * 64 + 64 + 2
    assert lib.gpc_stats_len(16, 8) % 2 == 0
    assert lib.gpc_omgp_padded_n(65) == 128 and lib.gpc_mohgp_eig_ws_len(64) == 64 + 4 * 64 * 64 + 8 and lib.gpc_mohgp_eig_ws_len(3) % 2 == 0


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    from gpclust_b200 import _cabi
    lib = _cabi.lib
    rc = lib.gpc_reduce_partials(None, 1, 1, None, None)
    assert rc == _cabi.GPC_EINVAL
    with pytest.raises(ValueError):
        _cabi.check(rc, "gpc_reduce_partials")
    with pytest.raises(ValueError):
        _cabi.padded_k(513)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gpclust_b200 import MOG, _cabi, utilities
    with pytest.raises(_cabi.GpclustLibraryError):
        MOG(np.random.randn(10, 2), K=2)
    with pytest.raises(_cabi.GpclustLibraryError):
        utilities.softmax(np.zeros((3, 2)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gpclust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
