"""Restarted, left-preconditioned GMRES on the device, following scipy.sparse.linalg.gmres.

Reference call site: /root/reference/code.py:516
    u, exit_code = scipy.sparse.linalg.gmres(A, f_vec, M=M, tol=1e-3, callback=counter_prec)
i.e. scipy's defaults restart=20, maxiter=10*N, x0=0, atol=0 and the legacy callback (called once per inner
iteration with the preconditioned residual estimate divided by ||b||).  The control flow below follows
scipy/sparse/linalg/_isolve/iterative.py (scipy 1.11+: modified Gram-Schmidt Arnoldi, Givens rotations,
inner tolerance adaptation ptol), so that iteration counts match the reference.

The n^2-sized work (matvec, preconditioner, dot products, axpys) runs in the CUDA kernels of
libhelmholtz_b200.so; only the (restart+1)^2 Hessenberg arithmetic is done on the host.  With a process
group, vectors are slab-distributed and the dot products are all-reduced (NCCL).
"""
import math

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def lartg(f, g):
    """Givens rotation with LAPACK zlartg semantics: real c, complex s, [c s; -conj(s) c] [f; g] = [r; 0]."""
    f, g = complex(f), complex(g)
    if g == 0:
        return 1.0, 0j, f
    if f == 0:
        d = abs(g)
        return 0.0, g.conjugate() / d, d
    f1 = abs(f)
    h = math.hypot(f1, abs(g))
    c = f1 / h
    fs = f / f1
    return c, fs * g.conjugate() / h, fs * h


class DeviceVectors:
    """Krylov vector kernels on (a slab of) the field; reductions are summed over `group` if given."""

    def __init__(self, nloc, device, group=None, restart=20):
        self.lib = _lib.require_device()
        self.nloc, self.device, self.group = nloc, torch.device(device), group
        self.scal = torch.zeros(max(64, restart + 3), dtype=torch.complex128, device=device)

    def reserve(self, restart):
        """hp_mgs writes restart + 2 scalars (coefficients, norm after, norm before)"""
        if self.scal.numel() < restart + 3:
            self.scal = torch.zeros(restart + 3, dtype=torch.complex128, device=self.device)

    def _reduce(self, t):
        if self.group is not None:
            import torch.distributed as dist
            r = torch.view_as_real(t)
            dist.all_reduce(r, group=self.group)
        return t

    def norm(self, x):
        if self.group is None:
            _lib.check(self.lib.hp_nrm2(self.nloc, _ptr(x), _ptr(self.scal), _stream()), "hp_nrm2")
            return float(self.scal[0].real.item())
        _lib.check(self.lib.hp_dotc(self.nloc, _ptr(x), _ptr(x), _ptr(self.scal), _stream()), "hp_dotc")
        return math.sqrt(float(self._reduce(self.scal[:1])[0].real.item()))

    def scale_copy(self, a, x, y):
        a = complex(a)
        _lib.check(self.lib.hp_scale_copy(self.nloc, a.real, a.imag, _ptr(x), _ptr(y), _stream()), "hp_scale_copy")

    def axpy(self, a, x, y):
        a = complex(a)
        _lib.check(self.lib.hp_axpy(self.nloc, a.real, a.imag, _ptr(x), _ptr(y), _stream()), "hp_axpy")

    def mgs(self, V, k, w):
        """Modified Gram-Schmidt of w against V[0..k): returns (h[0..k), ||w|| after, ||w|| before)."""
        if self.group is None:
            _lib.check(self.lib.hp_mgs(self.nloc, k, _ptr(V), V.stride(0), _ptr(w), _ptr(self.scal), _stream()), "hp_mgs")
            h = self.scal[:k + 2].cpu().numpy()
            return h[:k].copy(), float(h[k].real), float(h[k + 1].real)
        # distributed: the coefficients stay on the device (dot -> all-reduce -> axpy with a device scalar), one
        # copy to the host at the end
        import torch.distributed as dist
        sc = self.scal
        _lib.check(self.lib.hp_dotc(self.nloc, _ptr(w), _ptr(w), _ptr(sc[k + 1:]), _stream()), "hp_dotc")
        for j in range(k):
            _lib.check(self.lib.hp_dotc(self.nloc, _ptr(V[j]), _ptr(w), _ptr(sc[j:]), _stream()), "hp_dotc")
            dist.all_reduce(torch.view_as_real(sc[j:j + 1]), group=self.group)
            _lib.check(self.lib.hp_axpy_dev(self.nloc, _ptr(sc[j:]), -1.0, _ptr(V[j]), _ptr(w), _stream()), "hp_axpy_dev")
        _lib.check(self.lib.hp_dotc(self.nloc, _ptr(w), _ptr(w), _ptr(sc[k:]), _stream()), "hp_dotc")
        dist.all_reduce(torch.view_as_real(sc[k:k + 2]), group=self.group)
        h = sc[:k + 2].cpu().numpy()
        return h[:k].copy(), math.sqrt(float(h[k].real)), math.sqrt(float(h[k + 1].real))

    def combine(self, V, y, x):
        """x += sum_j y[j] V[j]."""
        y = np.ascontiguousarray(np.asarray(y, dtype=np.complex128))
        _lib.check(self.lib.hp_combine(self.nloc, len(y), _ptr(V), V.stride(0), y.ctypes.data, _ptr(x), _stream()),
                   "hp_combine")


def gmres(matvec, psolve, b, *, vec, rtol=1e-5, atol=0.0, restart=20, maxiter=None, callback=None, nglobal=None,
          health=None):
    """scipy.sparse.linalg.gmres(A, b, M=M, rtol=..., restart=..., maxiter=..., callback=...) on device vectors.

    matvec(x, out), psolve(x, out): device operators writing into `out`.  b: device vector (local slab).
    Returns (x, info, hist): hist holds what scipy hands to the legacy callback, one entry per inner iteration.
    """
    gen = gmres_steps(matvec, b, vec=vec, rtol=rtol, atol=atol, restart=restart, maxiter=maxiter, callback=callback,
                      nglobal=nglobal, health=health)
    try:
        req = next(gen)
        while True:
            psolve(*req)
            req = gen.send(None)
    except StopIteration as done:
        return done.value


def gmres_batch(matvec, psolve_batch, bs, *, vec, **kw):
    """The same iteration for several right-hand sides advanced in lock step: every round collects one preconditioner
    request (x, out) per unfinished system and hands the list to psolve_batch, which may pipeline them (slab.py sends
    them through the slabs one behind the other).  Returns [(x, info, hist), ...] in the order of `bs`."""
    gens = [gmres_steps(matvec, b, vec=vec, **kw) for b in bs]
    results = [None] * len(bs)
    reqs = {}
    for i, g in enumerate(gens):
        try:
            reqs[i] = next(g)
        except StopIteration as done:
            results[i] = done.value
    while reqs:
        order = sorted(reqs)
        psolve_batch([reqs[i] for i in order])
        for i in order:
            try:
                reqs[i] = gens[i].send(None)
            except StopIteration as done:
                results[i] = done.value
                del reqs[i]
    return results


def gmres_steps(matvec, b, *, vec, rtol=1e-5, atol=0.0, restart=20, maxiter=None, callback=None, nglobal=None,
                health=None):
    """Generator form of gmres(): yields (x, out) whenever the preconditioner has to be applied (out = M x) and
    returns (x, info, hist) through StopIteration.  health(): called at every restart boundary (the host is in sync
    with the device there anyway); raises if a kernel of the operators reported a fault."""
    nloc = b.numel()
    n = nglobal if nglobal is not None else nloc
    dev = b.device
    x = torch.zeros_like(b)
    hist = []
    bnrm2 = vec.norm(b)
    if bnrm2 == 0:
        return x, 0, hist
    atol = max(float(atol), float(rtol) * bnrm2)
    eps = np.finfo(np.float64).eps
    if maxiter is None:
        maxiter = n * 10
    restart = min(restart, n)
    vec.reserve(restart)
    V = torch.empty((restart + 1, nloc), dtype=torch.complex128, device=dev)
    r = torch.empty_like(b)
    av = torch.empty_like(b)
    w = torch.empty_like(b)
    yield (b, w)
    Mb_nrm2 = vec.norm(w)
    ptol_max_factor = 1.0
    ptol = Mb_nrm2 * min(ptol_max_factor, atol / bnrm2)
    presid = 0.0
    hh = np.zeros((restart, restart + 1), dtype=np.complex128)
    givens = np.zeros((restart, 2), dtype=np.complex128)
    inner_iter = 0
    rnorm = math.inf
    for iteration in range(maxiter):
        if iteration == 0:
            vec.scale_copy(1.0, b, r)
            if bnrm2 < atol:
                return x, 0, hist
        yield (r, V[0])
        tmp = vec.norm(V[0])
        vec.scale_copy(1.0 / tmp, V[0], V[0])
        S = np.zeros(restart + 1, dtype=np.complex128)
        S[0] = tmp
        breakdown = False
        col = 0
        for col in range(restart):
            matvec(V[col], av)
            yield (av, w)
            hcol, h1, h0 = vec.mgs(V, col + 1, w)
            hh[col, :col + 1] = hcol
            hh[col, col + 1] = h1
            if h1 <= eps * h0:
                hh[col, col + 1] = 0
                breakdown = True
                vec.scale_copy(1.0, w, V[col + 1])
            else:
                vec.scale_copy(1.0 / h1, w, V[col + 1])
            for k in range(col):
                c, s = givens[k, 0], givens[k, 1]
                n0, n1 = hh[col, k], hh[col, k + 1]
                hh[col, k], hh[col, k + 1] = c * n0 + s * n1, -np.conj(s) * n0 + c * n1
            c, s, mag = lartg(hh[col, col], hh[col, col + 1])
            givens[col, :] = [c, s]
            hh[col, col], hh[col, col + 1] = mag, 0
            tmp = -np.conjugate(s) * S[col]
            S[col], S[col + 1] = c * S[col], tmp
            presid = abs(tmp)
            inner_iter += 1
            hist.append(presid / bnrm2)
            if callback is not None:
                callback(presid / bnrm2)
            if inner_iter == maxiter:
                break
            if presid <= ptol or breakdown:
                break
        if hh[col, col] == 0:
            S[col] = 0
        y = np.zeros(col + 1, dtype=np.complex128)
        y[:] = S[:col + 1]
        for k in range(col, 0, -1):
            if y[k] != 0:
                y[k] /= hh[k, k]
                tmp = y[k]
                y[:k] -= tmp * hh[k, :k]
        if y[0] != 0:
            y[0] /= hh[0, 0]
        vec.combine(V, y, x)
        matvec(x, av)
        vec.scale_copy(-1.0, av, r)
        vec.axpy(1.0, b, r)
        rnorm = vec.norm(r)
        if health is not None:
            health()
        if inner_iter == maxiter:
            return x, (0 if rnorm <= atol else maxiter), hist
        if rnorm <= atol:
            break
        elif breakdown:
            break
        elif presid <= ptol:
            ptol_max_factor = max(eps, 0.25 * ptol_max_factor)
        else:
            ptol_max_factor = min(1.0, 1.5 * ptol_max_factor)
        ptol = presid * min(ptol_max_factor, atol / rnorm)
    info = 0 if rnorm <= atol else maxiter
    return x, info, hist
