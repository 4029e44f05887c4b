"""CPU check of the table-driven exp of the G pass (gpc_common.cuh: exp_batch_tab): the 32 table entries in the source are the
correctly rounded 2^(j/32), and the algorithm (Cody-Waite reduction by ln2/32, degree-6 Taylor polynomial, table entry, exponent
adjustment) restated with numpy stays within 2 ulp of a 64-bit-mantissa reference over the argument range of the path."""
import os
import re
import struct
from decimal import Decimal, getcontext

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table():
    src = open(os.path.join(ROOT, "gpclust_b200", "csrc", "gpc_common.cuh")).read()
    body = re.search(r"kExpTab\[32\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
    return [float.fromhex(t.strip()) for t in body.replace("\n", " ").split(",") if t.strip()]


def test_table_entries_are_correctly_rounded():
    tab = _table()
    assert len(tab) == 32
    getcontext().prec = 60
    for j, v in enumerate(tab):
        assert v == float(Decimal(2) ** (Decimal(j) / Decimal(32))), j


def test_algorithm_within_two_ulp():
    tab = _table()
    ln2hi, ln2lo, shift = 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6755399441055744.0
    fma = lambda a, b, c: np.float64(np.longdouble(a) * np.longdouble(b) + np.longdouble(c))
    coef = [1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0]
    rng = np.random.RandomState(0)
    xs = np.concatenate([rng.uniform(-60, 0, 4000), rng.uniform(-708, -650, 500), rng.uniform(-1e-3, 0, 500), [0.0, -708.0, 2.5, -0.010831]])
    worst = 0.0
    for x in xs:
        t = fma(x, 32 * 1.4426950408889634074, shift)
        n = t - shift
        ki = struct.unpack("<i", struct.pack("<d", t)[:4])[0]
        r = fma(n, -ln2lo / 32, fma(n, -ln2hi / 32, x))
        assert abs(r) <= np.log(2) / 64 * 1.001
        p = fma(1.0 / 720.0, r, coef[0])
        for c in coef[1:]:
            p = fma(p, r, c)
        v = tab[ki & 31] * p
        bits, = struct.unpack("<q", struct.pack("<d", v))
        v = struct.unpack("<d", struct.pack("<q", bits + ((ki >> 5) << 52)))[0]
        ref = np.exp(np.longdouble(x))
        worst = max(worst, float(abs(np.longdouble(v) - ref) / ref))
    assert worst < 2 * 2.220446049250313e-16, worst
