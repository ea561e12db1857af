"""csrc/pyset.cuh (the CPython set-order restatement the batched StrongSORT step runs on the device) against the running
interpreter: list(set(a) - set(b)) for the shapes linear_assignment.py:141 produces (a = confirmed track indices in
increasing order, b = matched ones in increasing order) and for arbitrary insertion orders."""
import ctypes as C
import os
import random
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if sys.version_info[:2] < (3, 8):
        pytest.skip("set implementation older than the one restated")
    so = str(tmp_path_factory.mktemp("pyset") / "pyset_check.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(HERE, "pyset_check.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.pyset_difference_order_c.restype = C.c_int
    lib.pyset_difference_order_c.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def run(lib, a, b, nkeys):
    aa, bb = np.asarray(a, dtype=np.int16), np.asarray(b, dtype=np.int16)
    out = np.zeros(max(len(a), 1), dtype=np.int16)
    n = lib.pyset_difference_order_c(aa.ctypes.data, len(aa), bb.ctypes.data, len(bb), nkeys, out.ctypes.data)
    return out[:n].tolist()


def test_matches_cpython_on_cascade_shapes(lib):
    rng = random.Random(5)
    for trial in range(3000):
        nkeys = rng.choice([1, 3, 8, 9, 20, 33, 64, 100, 129, 200, 256])
        na = rng.randint(0, nkeys)
        a = sorted(rng.sample(range(nkeys), na))
        frac = rng.choice([0.0, 0.05, 0.2, 0.24, 0.26, 0.5, 0.9, 1.0])
        b = sorted(rng.sample(a, int(round(frac * na))))
        ref = list(set(a) - set(k for k in b))
        assert run(lib, a, b, nkeys) == ref, (a, b)


def test_matches_cpython_on_arbitrary_orders(lib):
    rng = random.Random(7)
    for trial in range(2000):
        nkeys = rng.choice([5, 16, 40, 128, 256])
        a = rng.sample(range(nkeys), rng.randint(0, nkeys))
        b = rng.sample(range(nkeys), rng.randint(0, nkeys))      # b need not be a subset
        ref = list(set(a) - set(k for k in b))
        assert run(lib, a, b, nkeys) == ref, (a, b)
