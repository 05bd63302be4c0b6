"""Host-side mirror of the reference's entry points (/root/reference/code.py) over libhelmholtz_b200.so.

Same names, argument order and meaning as the reference for the hot path:

    build_A_matrix(b, const, eta, omega, h, n, c_mat)          code.py:202-219   -> DeviceCSR
    algo2_3(b, const, eta, omega, h, n, c_mat)                 code.py:345-353   -> (front, strips) handles
    algo2_4(f_vec, b, n, lu_HF, ..., lu_Hm_ra)                 code.py:356-385   -> M f
    run_solver(n, b, wave_num, const, alpha, init_func, ...)   code.py:424-541   -> SolveResult

PyTorch is used for device buffers and streams only; every n^2-sized operation is a kernel of the library.
There is no CPU path: without the library or a CUDA device these functions raise.
"""
import ctypes as C
import os
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, fields
from .gmres import DeviceVectors, gmres

DIAG_MODES = {"reference": 0, "paper": 1}
LAYOUT_MODES = {"auto": 0, "classic": 1, "cluster": 2}
FRONT_MODES = {"blockdiag": 0, "coupled": 1}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _as_device_field(x, device):
    """complex128 contiguous device tensor from numpy / torch input (host->device copy if needed)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.complex128).contiguous().reshape(-1)
    a = np.ascontiguousarray(np.asarray(x, dtype=np.complex128)).reshape(-1)
    return torch.from_numpy(a).to(device)


class HelmholtzSolver:
    """One problem instance on one device: owns the hp_solver handle (tables, velocity, factorisation).

    Arguments are the reference's: n interior points per side, b PML width in points, omega complex angular
    frequency (2*pi*wave_num + 1j*alpha, code.py:442), const the PML strength, c_mat the (n+2)^2 velocity."""

    def __init__(self, n, b, omega, const, c_mat, device=None):
        self.lib = _lib.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.n, self.b, self.omega, self.const = int(n), int(b), complex(omega), float(const)
        self.h = 1 / (n + 1)
        self.eta = b * self.h
        if isinstance(c_mat, torch.Tensor) and c_mat.is_cuda:
            c_dev = c_mat.to(dtype=torch.float64).contiguous()
            cptr, on_dev = c_dev.data_ptr(), 1
        else:
            c_host = np.ascontiguousarray(np.asarray(c_mat, dtype=np.float64))
            cptr, on_dev = c_host.ctypes.data, 0
        shape = tuple(c_mat.shape)
        if shape != (n + 2, n + 2):
            raise ValueError(f"c_mat must be ({n + 2}, {n + 2}) like the reference's init_c*_mat, got {shape}")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.hp_create(C.byref(h), self.n, self.b, self.omega.real, self.omega.imag, self.const,
                                          cptr, on_dev, _stream()), "hp_create")
        self.handle = h
        self.m_lo, self.m_hi = 0, -1
        self.max_group = 8            # cap on the right-hand sides per sweep launch (1 = one launch per vector)

    def clone_context(self):
        """A second handle on this solver's operator and factorisation with private sweep scratch (hp_context_clone):
        preconditioner applications issued through different contexts may be in flight on different CUDA streams at the
        same time (one context per group of right-hand sides in slab.GroupPipeline).  Close the contexts before the
        solver they were cloned from."""
        self._on_device()
        c = object.__new__(HelmholtzSolver)
        c.__dict__.update(self.__dict__)
        h = C.c_void_p()
        _lib.check(self.lib.hp_context_clone(self.handle, C.byref(h), _stream()), "hp_context_clone")
        c.handle = h
        c.parent = self                                      # keeps the owner of the factorisation alive
        return c

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- operator -------------------------------------------------------------------------------------
    def assemble_csr(self):
        """build_A_matrix: sorted CSR on the device (int32 indices, complex128 values)."""
        self._on_device()
        n = self.n
        N = n * n
        nnz = self.lib.hp_csr_nnz(n)
        indptr = torch.empty(N + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(nnz, dtype=torch.int32, device=self.device)
        data = torch.empty(nnz, dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.hp_assemble_csr(self.handle, _ptr(indptr), _ptr(indices), _ptr(data), _stream()),
                   "hp_assemble_csr")
        return DeviceCSR(self, indptr, indices, data)

    def assemble_strip_csr(self, m):
        """get_Hm(m): the strip operator of layer m as sorted CSR on the device (code.py:283-290)."""
        self._on_device()
        n, b = self.n, self.b
        nnz = self.lib.hp_strip_csr_nnz(n, b)
        indptr = torch.empty(b * n + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(nnz, dtype=torch.int32, device=self.device)
        data = torch.empty(nnz, dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.hp_assemble_strip_csr(self.handle, int(m), _ptr(indptr), _ptr(indices), _ptr(data), _stream()),
                   "hp_assemble_strip_csr")
        return DeviceCSR(self, indptr, indices, data, shape=(b * n, b * n))

    def matvec(self, x, out=None):
        """y = A x, matrix free."""
        self._on_device()
        if out is None:
            out = torch.empty_like(x)
        _lib.check(self.lib.hp_stencil_matvec(self.handle, _ptr(x), _ptr(out), _stream()), "hp_stencil_matvec")
        return out

    def matvec_rows(self, j_lo, j_hi, x, south, north, out):
        _lib.check(self.lib.hp_stencil_matvec_rows(self.handle, j_lo, j_hi, _ptr(x), _ptr(south), _ptr(north),
                                                   _ptr(out), _stream()), "hp_stencil_matvec_rows")
        return out

    # ---- preconditioner -------------------------------------------------------------------------------
    def setup_preconditioner(self, P=0, K=0, m_lo=0, m_hi=0, layout="auto", front="blockdiag"):
        """algo2_3.  (m_lo, m_hi) = (0, 0): all strips b+1..n; otherwise the strips of this rank's slab.
        layout: "auto" (cluster when a partition exists), "classic", "cluster" (include/helmholtz_b200.h).
        front: "blockdiag" = the reference's H_F (code.py:178-183, only the diagonal blocks), "coupled" = the full
        A[:bn, :bn] of the paper (2-5 GMRES iterations instead of 70-200 with diag='paper')."""
        self._on_device()
        _lib.check(self.lib.hp_set_front_mode(self.handle, FRONT_MODES[front]), "hp_set_front_mode")
        self.front = front
        _lib.check(self.lib.hp_set_layout_mode(self.handle, LAYOUT_MODES[layout]), "hp_set_layout_mode")
        _lib.check(self.lib.hp_precond_setup(self.handle, P, K, m_lo, m_hi, _stream()), "hp_precond_setup")
        if m_lo == 0 and m_hi == 0:
            m_lo, m_hi = self.b + 1, self.n
        self.m_lo, self.m_hi = m_lo, m_hi
        return self

    def set_front(self, front):
        """switch the front block of a solver that is set up ("blockdiag" | "coupled"); the strips are kept"""
        self._on_device()
        _lib.check(self.lib.hp_precond_set_front(self.handle, FRONT_MODES[front], _stream()), "hp_precond_set_front")
        self.front = front
        return self

    def set_sweep_variant(self, variant):
        """0 = automatic; classic layout: 1 direct, 2 TMA staged, 3 pipelined; cluster layout: 4."""
        _lib.check(self.lib.hp_set_sweep_variant(self.handle, int(variant)), "hp_set_sweep_variant")
        return self

    def sweep_status(self):
        """0 = fine (synchronises the device)."""
        return int(self.lib.hp_sweep_status(self.handle))

    def check_status(self):
        """Raise if a sweep kernel gave up waiting for another CTA since the setup (its output is then garbage).
        Synchronises the device; called where the host waits for the device anyway (GMRES restart boundaries, the end
        of run_solver, algo2_4)."""
        st = self.sweep_status()
        if st:
            raise _lib.HelmholtzB200Error(
                f"sweep kernel fault (status {st}): a CTA timed out waiting for exchange data; results since the last "
                "setup_preconditioner() are invalid")

    def _on_device(self):
        """every call runs on the current device and stream: refuse a solver that lives elsewhere"""
        if torch.cuda.current_device() != self.device.index:
            raise _lib.HelmholtzB200Error(
                f"solver lives on {self.device}, current device is cuda:{torch.cuda.current_device()} "
                "(wrap the call in torch.cuda.device(solver.device))")

    @property
    def precond_bytes(self):
        return int(self.lib.hp_precond_bytes(self.handle))

    @property
    def setup_ms(self):
        return float(self.lib.hp_precond_setup_ms(self.handle))

    def layout(self):
        P, K, QP, CW, NS, NR = (C.c_int() for _ in range(6))
        PK = C.c_int64()
        _lib.check(self.lib.hp_strip_layout(self.handle, C.byref(P), C.byref(K), C.byref(QP), C.byref(CW), C.byref(NS),
                                            C.byref(NR), C.byref(PK), None, None, None), "hp_strip_layout")
        ls = (C.c_int * P.value)()
        lq = (C.c_int * P.value)()
        sp = (C.c_int * max(P.value - 1, 1))()
        _lib.check(self.lib.hp_strip_layout(self.handle, None, None, None, None, None, None, None, ls, lq, sp),
                   "hp_strip_layout")
        colN, NCB, NRQ, NXG = (C.c_int() for _ in range(4))
        _lib.check(self.lib.hp_strip_layout_ex(self.handle, C.byref(colN), C.byref(NCB), C.byref(NRQ), C.byref(NXG)),
                   "hp_strip_layout_ex")
        return dict(P=P.value, K=K.value, G=P.value * K.value, QP=QP.value, CW=CW.value, NS=NS.value, NR=NR.value,
                    PK=PK.value, leaf_start=np.array(ls[:]), q=np.array(lq[:]), sep=np.array(sp[:P.value - 1]),
                    colN=colN.value, NCB=NCB.value, NRQ=NRQ.value, NXG=NXG.value)

    def strip_packets(self, m):
        L = self.layout()
        out = np.zeros((L["G"], L["PK"]), dtype=np.complex128)
        _lib.check(self.lib.hp_strip_packets(self.handle, m, out.ctypes.data), "hp_strip_packets")
        return L, out

    def strip_apply(self, m, v, out=None):
        """T_m v: last n entries of H_m^{-1} [0; v]  (code.py:368-370)."""
        self._on_device()
        if out is None:
            out = torch.empty_like(v)
        _lib.check(self.lib.hp_strip_apply(self.handle, m, _ptr(v), _ptr(out), _stream()), "hp_strip_apply")
        return out

    def precond_apply(self, f, out=None, diag="reference"):
        """algo2_4: out = M f."""
        self._on_device()
        if out is None:
            out = torch.empty_like(f)
        _lib.check(self.lib.hp_precond_apply(self.handle, _ptr(f), _ptr(out), DIAG_MODES[diag], _stream()),
                   "hp_precond_apply")
        return out

    @property
    def multi_max(self):
        """right-hand sides one sweep launch can carry with the partition in use (1, 2, 4 or 8)"""
        return int(self.lib.hp_multi_max(self.handle))

    def batch_group(self, R):
        """right-hand sides one sweep launch carries when R are in flight"""
        g = 1
        while g * 2 <= min(R, self.multi_max, self.max_group):
            g *= 2
        return g

    def batch_kernel_name(self, R):
        g = self.batch_group(R)
        if g > 1:
            return "hp_sweep4d_kernel (FP64 tensor cores, 8 right-hand sides)" if g == 8 and not os.environ.get("HP_NO_DMMA") else "hp_sweep4m_kernel"
        return "hp_sweep4_kernel" if self.layout()["colN"] else "hp_sweep2_kernel"

    @staticmethod
    def _ptr_array(ts):
        return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

    def precond_apply_multi(self, fs, outs, diag="reference"):
        """algo2_4 on R = len(fs) vectors with one pass over the strip generators per sweep (R in {1, 2, 4, 8})."""
        self._on_device()
        _lib.check(self.lib.hp_precond_apply_multi(self.handle, len(fs), self._ptr_array(fs), self._ptr_array(outs), DIAG_MODES[diag],
                                                   _stream()), "hp_precond_apply_multi")
        return outs

    def precond_apply_batch(self, pairs, diag="reference"):
        """out_i = M x_i for every (x_i, out_i) of `pairs`, in groups of batch_group() right-hand sides per launch"""
        i = 0
        while i < len(pairs):
            g = self.batch_group(len(pairs) - i)
            if g == 1:
                self.precond_apply(pairs[i][0], out=pairs[i][1], diag=diag)
            else:
                self.precond_apply_multi([p[0] for p in pairs[i:i + g]], [p[1] for p in pairs[i:i + g]], diag=diag)
            i += g

    def sweep_forward_multi_buf(self, bufs, row0, m_from, m_to):
        arr = (C.c_void_p * len(bufs))(*[self._base(t, row0) for t in bufs])
        _lib.check(self.lib.hp_sweep_forward_multi(self.handle, len(bufs), arr, m_from, m_to, _stream()), "hp_sweep_forward_multi")

    def sweep_backward_multi_buf(self, bufs, row0, m_from, m_to, diag="reference"):
        arr = (C.c_void_p * len(bufs))(*[self._base(t, row0) for t in bufs])
        _lib.check(self.lib.hp_sweep_backward_multi(self.handle, len(bufs), arr, m_from, m_to, DIAG_MODES[diag], _stream()),
                   "hp_sweep_backward_multi")

    # staged calls on a slab buffer whose first row is global row `row0` (slab decomposition, slab.py): the
    # kernels index the field by absolute row, so they get the address global row 0 would have
    def _base(self, buf, row0):
        return buf.data_ptr() - row0 * self.n * 16

    def front_begin_buf(self, buf, row0):
        _lib.check(self.lib.hp_front_begin(self.handle, self._base(buf, row0), _stream()), "hp_front_begin")

    def front_end_buf(self, buf, row0):
        _lib.check(self.lib.hp_front_end(self.handle, self._base(buf, row0), _stream()), "hp_front_end")

    def front_tf_new(self):
        """buffer that can park T_F u_F of one right-hand side (front_tf_save / front_tf_load)"""
        return torch.empty(self.b * self.n, dtype=torch.complex128, device=self.device)

    def front_tf_save(self, t):
        _lib.check(self.lib.hp_front_tf_copy(self.handle, _ptr(t), 0, _stream()), "hp_front_tf_copy")

    def front_tf_load(self, t):
        _lib.check(self.lib.hp_front_tf_copy(self.handle, _ptr(t), 1, _stream()), "hp_front_tf_copy")

    def sweep_forward_buf(self, buf, row0, m_from, m_to):
        _lib.check(self.lib.hp_sweep_forward(self.handle, self._base(buf, row0), m_from, m_to, _stream()), "hp_sweep_forward")

    def sweep_backward_buf(self, buf, row0, m_from, m_to, diag="reference"):
        _lib.check(self.lib.hp_sweep_backward(self.handle, self._base(buf, row0), m_from, m_to, DIAG_MODES[diag], _stream()),
                   "hp_sweep_backward")

    # staged calls (slab decomposition)
    def front_begin(self, u):
        _lib.check(self.lib.hp_front_begin(self.handle, _ptr(u), _stream()), "hp_front_begin")

    def front_end(self, u):
        _lib.check(self.lib.hp_front_end(self.handle, _ptr(u), _stream()), "hp_front_end")

    def sweep_forward(self, u, m_from, m_to):
        _lib.check(self.lib.hp_sweep_forward(self.handle, _ptr(u), m_from, m_to, _stream()), "hp_sweep_forward")

    def sweep_backward(self, u, m_from, m_to, diag="reference"):
        _lib.check(self.lib.hp_sweep_backward(self.handle, _ptr(u), m_from, m_to, DIAG_MODES[diag], _stream()),
                   "hp_sweep_backward")


class DeviceCSR:
    """What build_A_matrix returns: the operator as sorted CSR arrays in device memory."""

    def __init__(self, solver, indptr, indices, data, shape=None):
        self.solver, self.indptr, self.indices, self.data = solver, indptr, indices, data
        self.shape = shape if shape is not None else (solver.n ** 2, solver.n ** 2)

    def matvec(self, x, out=None):
        x = _as_device_field(x, self.solver.device)
        if out is None:
            out = torch.empty_like(x)
        _lib.check(self.solver.lib.hp_csr_matvec(self.shape[0], _ptr(self.indptr), _ptr(self.indices), _ptr(self.data),
                                                 _ptr(x), _ptr(out), _stream()), "hp_csr_matvec")
        return out

    __matmul__ = matvec

    def to_host(self):
        """(indptr, indices, data) as numpy arrays (scipy.sparse.csr_matrix takes them as is)."""
        return self.indptr.cpu().numpy(), self.indices.cpu().numpy(), self.data.cpu().numpy()


# ------------------------------------------------------------------------------------------------------
# reference-named functions
# ------------------------------------------------------------------------------------------------------
def _solver_for(b, const, eta, omega, h, n, c_mat, device=None):
    if abs(h - 1 / (n + 1)) > 1e-15 or abs(eta - b * h) > 1e-15:
        raise ValueError("h and eta must be 1/(n+1) and b*h as in the reference (code.py:443-444)")
    return HelmholtzSolver(n, b, omega, const, c_mat, device=device)


def build_A_matrix(b, const, eta, omega, h, n, c_mat, device=None):
    """code.py:202-219."""
    return _solver_for(b, const, eta, omega, h, n, c_mat, device).assemble_csr()


def get_Hm(m, b, const, eta, omega, h, n, c_mat, device=None):
    """code.py:283-290: the strip operator of layer m (DeviceCSR, bn x bn)."""
    return _solver_for(b, const, eta, omega, h, n, c_mat, device).assemble_strip_csr(m)


def get_A_FF_block(b, const, eta, omega, h, n, c_mat, device=None, coupled=False):
    """code.py:178-183.  The reference keeps only the diagonal blocks (b tridiagonal systems); coupled=True is the full
    A[:bn, :bn] = get_Hm(b).  Returned as DeviceCSR; the block-diagonal form drops the couplings between the rows."""
    A = _solver_for(b, const, eta, omega, h, n, c_mat, device).assemble_strip_csr(b)
    if coupled:
        return A
    indptr, indices, data = A.to_host()
    rows = np.repeat(np.arange(b * n), np.diff(indptr))
    keep = np.abs(indices - rows) <= 1                              # code.py:180-182: block_diag of the A_ii
    ip = np.zeros(b * n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows[keep], minlength=b * n), out=ip[1:])
    dev = A.solver.device
    return DeviceCSR(A.solver, torch.from_numpy(ip).to(dev), torch.from_numpy(indices[keep]).to(dev),
                     torch.from_numpy(data[keep]).to(dev), shape=A.shape)


def algo2_3(b, const, eta, omega, h, n, c_mat, device=None, P=0, K=0, layout="auto", front="blockdiag"):
    """code.py:345-353.  Returns (lu_HF, lu_Hm_ra): both are the same factorisation handle here."""
    s = _solver_for(b, const, eta, omega, h, n, c_mat, device).setup_preconditioner(P, K, layout=layout, front=front)
    return s, s


def algo2_4(f_vec, b, n, lu_HF, A_b1F=None, A_Fb1=None, up_A_ra=None, lo_A_ra=None, lu_Hm_ra=None, diag="reference"):
    """code.py:356-385.  The coupling blocks A_b1F, A_Fb1, up_A_ra, lo_A_ra of the reference signature are
    implied by the solver handle and ignored.  Returns M f as an (n, n) device tensor like the reference's u."""
    s = lu_HF
    f = _as_device_field(f_vec, s.device)
    u = s.precond_apply(f, diag=diag).reshape(n, n)
    s.check_status()
    return u


@dataclass
class SolveResult:
    u: torch.Tensor                 # solution field, (n*n,) complex128 on the device (reference: u)
    residuals: list                 # what scipy hands to the callback, one per inner iteration
    niter: int                      # counter_prec.niter of the reference
    info: int                       # exit_code of scipy gmres
    init_time: float                # "Initialization time" (code.py:522)
    solve_time: float               # "GMRES solve time"    (code.py:523)
    solver: object = field(default=None, repr=False)

    def __iter__(self):             # the reference returns (init_time_length, solve_time_length)
        return iter((self.init_time, self.solve_time))


def run_solver(n, b, wave_num, const, alpha, init_func=fields.init_c1_f1, plot_solution=False, *, c_mat=None,
               f_mat=None, diag="reference", precond_input="rhs", front="blockdiag", rtol=1e-3, restart=20, maxiter=None,
               device=None, P=0, K=0, verbose=True, solver=None):
    """code.py:424-541 (the preconditioned solve; plotting is not part of this package).

    precond_input='rhs' is the reference as written: its LinearOperator ignores the vector it is given and
    always returns algo2_4(f_vec) (code.py:510-511).  'vector' applies the preconditioner to the argument.
    diag='reference' keeps u_m <- u_m - T_m u_m (code.py:372-375); 'paper' is Engquist-Ying's u_m <- T_m u_m.
    front='blockdiag' is the reference's H_F (code.py:178-183); 'coupled' keeps the couplings between its rows.
    The reference as written (rhs / reference / blockdiag) does not converge (info != 0); vector / paper converges,
    with front='coupled' in a handful of iterations.  A solver passed in keeps the front block it was set up with.
    """
    lib = _lib.require_device()   # noqa: F841  (fail before any host work if the device path is missing)
    t0 = time.time()
    omega = 2 * np.pi * wave_num + 1j * alpha          # code.py:442
    if c_mat is None or f_mat is None:
        c0, f0 = init_func(omega, n)
        c_mat = c0 if c_mat is None else c_mat
        f_mat = f0 if f_mat is None else f_mat
    s = solver if solver is not None else HelmholtzSolver(n, b, omega, const, c_mat, device=device)
    f = _as_device_field(f_mat, s.device)                # f_mat.flatten(), code.py:448
    if solver is None:
        s.setup_preconditioner(P, K, front=front)        # algo2_3, code.py:496
    torch.cuda.synchronize(s.device)
    t1 = time.time()
    vec = DeviceVectors(n * n, s.device, restart=restart)
    Mf = None
    if precond_input == "rhs":
        Mf = s.precond_apply(f, diag=diag)

        def psolve(x, out):
            # code.py:510-511: matvec=lambda x: algo2_4(f_vec, ...) -- the argument is ignored; algo2_4 is
            # deterministic, so its result is computed once and copied
            vec.scale_copy(1.0, Mf, out)
    elif precond_input == "vector":
        def psolve(x, out):
            s.precond_apply(x, out=out, diag=diag)
    else:
        raise ValueError("precond_input must be 'rhs' or 'vector'")
    u, info, hist = gmres(lambda x, out: s.matvec(x, out), psolve, f, vec=vec, rtol=rtol, restart=restart,
                          maxiter=maxiter, health=s.check_status)
    s.check_status()
    t2 = time.time()
    if verbose:
        print("GMRES iterations with preconditioner: " + str(len(hist)))      # code.py:520
        print("Initialization time = " + str(t1 - t0))
        print("GMRES solve time = " + str(t2 - t1))
    return SolveResult(u, hist, len(hist), info, t1 - t0, t2 - t1, s)
