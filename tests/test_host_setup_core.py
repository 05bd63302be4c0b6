"""csrc/hp_small.h + csrc/hp_setup_core.h compiled for the CPU (tests/host_harness.cpp) against the numpy
model tools/strip_model.py and the oracle.  These are the exact per-thread bodies the setup kernels run."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import helmholtz_oracle as orc
from tools import strip_model as sm

HERE = os.path.dirname(os.path.abspath(__file__))
C = ctypes
cp = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def hh():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libhost_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_harness.cpp")])
    lib = C.CDLL(so)
    lib.hh_tables.argtypes = [C.c_int] + [C.c_double] * 5 + [cp] * 4
    lib.hh_inv.argtypes = [C.c_int, cp]
    lib.hh_strip_setup.argtypes = [C.c_int] * 5 + [ip, ip, ip] + [C.c_double] * 5 + [dp, cp, cp, cp]
    return lib


def problem(n, b, wn, const):
    omega = 2 * np.pi * wn + 2j
    h = 1 / (n + 1)
    return dict(b=b, const=const, eta=b * h, omega=omega, h=h, n=n), orc.init_c1_f1(omega, n)[0]


def rel(a, b):
    d = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (d if d > 0 else 1.0)


def test_inverse(hh):
    rng = np.random.default_rng(0)
    for b in (1, 2, 5, 12, 16, 20, 24):
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


def strip_setup(hh, p, c_mat, m, P):
    n, b = p["n"], p["b"]
    pt = sm.partition(n, P, 1)
    QP, ns = pt["QP"], P - 1
    W = np.zeros((P, QP, QP), np.complex128)
    G = np.zeros((P, 2, b, QP), np.complex128)
    N = np.zeros((max(ns * b, 1), max(ns * b, 1)), np.complex128)
    bad = hh.hh_strip_setup(n, b, m, P, QP, pt["leaf_start"].astype(np.int32), pt["q"].astype(np.int32),
                            np.ascontiguousarray(pt["sep"].astype(np.int32)) if ns else np.zeros(1, np.int32),
                            p["const"], p["eta"], p["h"], p["omega"].real, p["omega"].imag,
                            np.ascontiguousarray(c_mat), W, G, N)
    assert bad == 0
    return pt, W, G, N


@pytest.mark.parametrize("n,b,wn,const,P", [(45, 12, 6, 70, 4), (63, 12, 4, 61, 5), (40, 5, 4, 30, 1),
                                            (40, 5, 4, 30, 7), (50, 20, 5, 60, 3)])
def test_strip_generators(hh, n, b, wn, const, P):
    p, c_mat = problem(n, b, wn, const)
    for m in (b + 1, (n + b) // 2, n):
        pt, W, G, N = strip_setup(hh, p, c_mat, m, P)
        mod = sm.StripModel(m, c_mat=c_mat, P=P, **p)
        assert rel(W, mod.W) < 1e-11
        assert rel(G, mod.G) < 1e-11
        if P > 1:
            assert rel(N, mod.N) < 1e-11


def test_strip_apply_medium(hh):
    """n = 200: T_m v from the C++ generators against the oracle's splu solve (code.py:368-370)."""
    n, b, P = 200, 12, 6
    p, c_mat = problem(n, b, 20, 80)
    Pc = orc.SweepingPreconditioner(c_mat=c_mat, **p)
    rng = np.random.default_rng(1)
    for m in (b + 1, 117, n):
        pt, W, G, N = strip_setup(hh, p, c_mat, m, P)
        mod = sm.StripModel.__new__(sm.StripModel)
        mod.b, mod.n, mod.m, mod.part, mod.W, mod.G, mod.N = b, n, m, pt, W, G, N
        v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert rel(mod.apply(v), Pc.T(m, v)) < 1e-12


@pytest.mark.parametrize("seed", range(6))
def test_half_warp_inverse_model(seed):
    """The rolled, register-rotating Gauss-Jordan of the b = 12 setup kernels (hp_half_inv, csrc/hp_setup.cu) as a lane-level
    numpy model: blocks that need no row exchange, blocks that need one at every step, a Schur block of a strip."""
    rng = np.random.default_rng(seed)
    B = 12
    if seed < 2:                                     # diagonally dominant: the pivot word is the identity permutation
        A = rng.standard_normal((B, B)) + 1j * rng.standard_normal((B, B)) + 40 * np.eye(B)
    elif seed < 4:                                   # dominant anti-diagonal: a row exchange at (almost) every step
        A = rng.standard_normal((B, B)) + 1j * rng.standard_normal((B, B)) + 40 * np.fliplr(np.eye(B))
    else:                                            # diagonal block of a strip operator (tridiagonal, complex symmetric)
        n, b = 40, B
        omega = 2 * np.pi * 4 + 2j
        h = 1 / (n + 1)
        c_mat = orc.init_c1_f1(omega, n)[0]
        D, L, U = sm.strip_blocks(b + 5 + seed, b, 60.0, b * h, omega, h, n, c_mat)
        A = np.asarray(D[n // 2])
    Ai, pivs = sm.half_warp_inverse_model(A)
    assert np.linalg.norm(Ai @ A - np.eye(B)) < 1e-12 * np.linalg.cond(A)
    assert np.allclose(Ai, np.linalg.inv(A), rtol=1e-10, atol=1e-12 * np.abs(np.linalg.inv(A)).max())
    if seed < 2:
        assert pivs == sum(p << (4 * p) for p in range(B))
    elif seed < 4:
        assert pivs != sum(p << (4 * p) for p in range(B))
