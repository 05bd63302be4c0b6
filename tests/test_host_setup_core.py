"""csrc/hp_small.h + csrc/hp_setup_core.h compiled for the CPU (tests/host_harness.cpp) against the numpy
model tools/tree_prototype.py.  These are the exact functions the setup kernels call per thread."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import helmholtz_oracle as orc
from tools import tree_prototype as tp

HERE = os.path.dirname(os.path.abspath(__file__))
C = ctypes
cp = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def hh():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libhost_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_harness.cpp")])
    lib = C.CDLL(so)
    lib.hh_tables.argtypes = [C.c_int] + [C.c_double] * 5 + [cp] * 4
    lib.hh_leaf_chains.argtypes = [C.c_int] * 5 + [C.c_double] * 5 + [dp, cp, cp, cp]
    lib.hh_merge.argtypes = [C.c_int, cp, cp, cp, cp, cp]
    lib.hh_coupling.argtypes = [C.c_int] * 4 + [C.c_double] * 5 + [cp]
    lib.hh_inv.argtypes = [C.c_int, cp]
    return lib


def problem(n, b, wn, const):
    omega = 2 * np.pi * wn + 2j
    h = 1 / (n + 1)
    return dict(b=b, const=const, eta=b * h, omega=omega, h=h, n=n), orc.init_c1_f1(omega, n)[0]


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def test_inverse(hh):
    rng = np.random.default_rng(0)
    for b in (1, 2, 5, 12, 16):
        A = rng.standard_normal((b, b)) + 1j * rng.standard_normal((b, b))
        if b > 1:
            A[0, 0] = 0  # force a row exchange
        X = np.ascontiguousarray(A.copy())
        assert hh.hh_inv(b, X) == 0
        assert rel(X, np.linalg.inv(A)) < 1e-12


def test_tables(hh):
    p, _ = problem(45, 12, 6, 70)
    n = p["n"]
    out = [np.zeros(2 * n + 3, np.complex128) for _ in range(4)]
    hh.hh_tables(n, p["const"], p["eta"], p["h"], p["omega"].real, p["omega"].imag, *out)
    x = np.arange(2 * n + 3) * 0.5 * p["h"]
    assert rel(out[0], orc.s1(x, p["const"], p["eta"], p["omega"])) < 1e-15
    assert rel(out[2], orc.s2(x, p["const"], p["eta"], p["omega"])) < 1e-15
    assert rel(out[1] * out[0], np.ones_like(x)) < 1e-15


@pytest.mark.parametrize("n,b,wn,const,qmax", [(45, 12, 6, 70, 8), (63, 12, 4, 61, 16), (40, 5, 4, 30, 64)])
def test_chains_and_merge(hh, n, b, wn, const, qmax):
    p, c_mat = problem(n, b, wn, const)
    c_mat = np.ascontiguousarray(c_mat)
    for m in (b + 1, (n + b) // 2, n):
        tree = tp.StripTree(m, c_mat=c_mat, qmax=qmax, **p)
        st = tree.start
        Finv = np.zeros((n, b * b), np.complex128)
        Binv = np.zeros((n, b * b), np.complex128)
        gcol = np.zeros((n, b), np.complex128)
        for l in range(tree.P):
            bad = hh.hh_leaf_chains(n, b, m, int(st[l]) + 1, int(st[l + 1]), p["const"], p["eta"], p["h"],
                                    p["omega"].real, p["omega"].imag, c_mat, Finv, Binv, gcol)
            assert bad == 0
            lf = tree.leaves[l]
            sl = slice(st[l], st[l + 1])
            assert rel(Finv[sl], lf["Finv"].reshape(-1, b * b)) < 1e-12
            assert rel(Binv[sl], lf["Binv"].reshape(-1, b * b)) < 1e-12
            assert rel(gcol[sl], lf["gcol"]) < 1e-12
        # merges, level by level, fed with the prototype's corners
        for lv in range(1, tree.d + 1):
            for t in range(tree.P >> lv):
                c1, c2 = tree.corners[lv - 1][2 * t], tree.corners[lv - 1][2 * t + 1]
                pack = lambda c: np.ascontiguousarray(np.stack([c["pp"], c["pt"], c["tp"], c["tt"]]))  # noqa: E731
                q1 = int(st[(2 * t + 1) << (lv - 1)])            # 0-based first row of right child = q (1-based)
                cpl = np.zeros(b, np.complex128)
                hh.hh_coupling(n, b, m, q1, p["const"], p["eta"], p["h"], p["omega"].real, p["omega"].imag, cpl)
                assert rel(cpl, tree.U[q1 - 1]) < 1e-14 and rel(cpl, tree.L[q1]) < 1e-14
                rec = np.zeros(12 * b * b, np.complex128)
                corners = np.zeros((4, b, b), np.complex128)
                assert hh.hh_merge(b, pack(c1), pack(c2), cpl, rec, corners) == 0
                assert rel(rec, tree.nodes[tree.lvoff[lv] + t]) < 1e-11
                me = tree.corners[lv][t]
                assert rel(corners, np.stack([me["pp"], me["pt"], me["tp"], me["tt"]])) < 1e-11
