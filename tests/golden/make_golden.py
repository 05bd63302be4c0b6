"""Generate golden vectors from the UNMODIFIED reference (/root/reference/code.py).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported as a module (numba enabled).  Two shims are needed
for it to import/run on this image and neither touches its arithmetic:
  * matplotlib is absent  -> an empty stub module is registered;
  * scipy 1.18 removed gmres' ``tol=`` keyword (code.py:516 uses it) -> a
    wrapper maps ``tol`` to ``rtol``.
Everything stored here is the output of the reference's own functions.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference(path="/root/reference/code.py"):
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    orig = spla.gmres
    if not getattr(orig, "_tol_shim", False):
        def gmres(A, b, x0=None, tol=None, **kw):
            if tol is not None:
                kw["rtol"] = tol
            return orig(A, b, x0, **kw)
        gmres._tol_shim = True
        spla.gmres = gmres
    spec = importlib.util.spec_from_file_location("ref_code", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def one_case(ref, name, n, b, wave_num, const, alpha, init_name, with_gmres=True):
    init = getattr(ref, init_name)
    omega = 2 * np.pi * wave_num + 1j * alpha
    h = 1 / (n + 1)
    eta = b * h
    c_mat, f_mat = init(omega, n)
    f_vec = f_mat.flatten()
    A = ref.build_A_matrix(b, const, eta, omega, h, n, c_mat).tocsr()
    A.sort_indices()
    out = dict(n=n, b=b, wave_num=wave_num, const=const, alpha=alpha, init=init_name,
               c_mat=c_mat, f_mat=f_mat,
               A_indptr=A.indptr.astype(np.int32), A_indices=A.indices.astype(np.int32), A_data=A.data)
    rng = np.random.default_rng(1234)
    xr = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    out["x_rand"] = xr
    out["A_x_rand"] = A @ xr
    # strip operators of the first, a middle and the last moving-PML layer (code.py:283-290)
    for tag, m in (("first", b + 1), ("mid", (b + 1 + n) // 2), ("last", n)):
        Hm = ref.get_Hm(m, b, const, eta, omega, h, n, c_mat).tocsr()
        Hm.sort_indices()
        out[f"Hm_{tag}_m"] = m
        out[f"Hm_{tag}_indptr"] = Hm.indptr.astype(np.int32)
        out[f"Hm_{tag}_indices"] = Hm.indices.astype(np.int32)
        out[f"Hm_{tag}_data"] = Hm.data
    if with_gmres:
        lu_HF, lu_Hm_ra = ref.algo2_3(b, const, eta, omega, h, n, c_mat)
        A_b1F = ref.get_A_b1F_block(b, const, eta, omega, h, n, c_mat)
        A_Fb1 = ref.get_A_Fb1_block(b, const, eta, omega, h, n, c_mat)
        up_A_ra, lo_A_ra = [], []
        for i in range(1, n):
            up_A_ra.append(ref.get_A_block(i, i + 1, b, const, eta, omega, h, n, c_mat))
            lo_A_ra.append(ref.get_A_block(i + 1, i, b, const, eta, omega, h, n, c_mat))
        args = (b, n, lu_HF, A_b1F, A_Fb1, up_A_ra, lo_A_ra, lu_Hm_ra)
        out["M_f"] = np.asarray(ref.algo2_4(f_vec, *args)).ravel()           # what code.py:510 computes
        out["M_x_rand"] = np.asarray(ref.algo2_4(xr, *args)).ravel()         # algo2_4 on another vector
        # T_m v for the middle layer: last n entries of Hm^{-1} [0; v]
        m = out["Hm_mid_m"]
        t = np.zeros(b * n, dtype=np.cdouble)
        t[-n:] = xr[:n]
        out["T_mid_v"] = lu_Hm_ra[m - b - 1].solve(t)[-n:]
        # the literal solve of code.py:510-516
        hist = []
        M = spla.LinearOperator((n * n, n * n), matvec=lambda x: ref.algo2_4(f_vec, *args))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            u, info = spla.gmres(A, f_vec, M=M, tol=1e-3, callback=lambda r: hist.append(r))
        out["gmres_literal_u"] = u
        out["gmres_literal_info"] = info
        out["gmres_literal_hist"] = np.array(hist)
        # same reference operators, but M applied to the vector it is given (bounded)
        hist2 = []
        M2 = spla.LinearOperator((n * n, n * n), matvec=lambda x: ref.algo2_4(x, *args))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            u2, info2 = spla.gmres(A, f_vec, M=M2, tol=1e-3, maxiter=25, callback=lambda r: hist2.append(r))
        out["gmres_vector_u"] = u2
        out["gmres_vector_info"] = info2
        out["gmres_vector_hist"] = np.array(hist2)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("A_", "gmres"))})


if __name__ == "__main__":
    ref = load_reference()
    # small case with every init function (assembly + source only)
    one_case(ref, "ref_n20_b5_c1f1", 20, 5, 3, 30, 2, "init_c1_f1")
    one_case(ref, "ref_n20_b5_c1f2", 20, 5, 3, 30, 2, "init_c1_f2", with_gmres=False)
    one_case(ref, "ref_n20_b5_c2f1", 20, 5, 3, 30, 2, "init_c2_f1", with_gmres=False)
    one_case(ref, "ref_n20_b5_c2f2", 20, 5, 3, 30, 2, "init_c2_f2", with_gmres=False)
    # the reference's own commented example, code.py:570  run_solver(63, 12, 4, 61, 2, init_c1_f1)
    one_case(ref, "ref_n63_b12_c1f1", 63, 12, 4, 61, 2, "init_c1_f1")
    # non-power-of-two interior size, heterogeneous c2 with plane-wave-modulated source
    one_case(ref, "ref_n45_b12_c2f2", 45, 12, 6, 70, 2, "init_c2_f2")
