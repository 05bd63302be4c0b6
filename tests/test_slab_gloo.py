"""Host logic of the slab decomposition (helmholtz_preconditioner_b200/slab.py) with world_size 2 and 3 over gloo on
the CPU: the message schedule, halo exchange and strip ownership are driven with a numpy stand-in for the
per-rank kernels (built on the oracle), and the distributed result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import helmholtz_oracle as orc


class NumpyBackend:
    """The staged per-rank operations of HelmholtzSolver, on CPU tensors, from the oracle's pieces."""

    def __init__(self, n, b, omega, const, c_mat):
        h = 1 / (n + 1)
        self.n, self.b = n, b
        self.p = dict(b=b, const=const, eta=b * h, omega=omega, h=h, n=n, c_mat=c_mat)
        self.P = orc.SweepingPreconditioner(**self.p)
        self.c = orc.stencil_coeffs(np.arange(1, n + 1), None, **self.p)

    def _rows(self, buf, row0):
        return buf.numpy().reshape(-1, self.n), row0

    def front_begin_buf(self, buf, row0):
        u, r0 = self._rows(buf, row0)
        b, n = self.b, self.n
        self.TF = self.P.lu_HF.solve(u[0 - r0:b - r0].ravel())
        u[b - r0] -= self.P.lo[b] * self.TF[-n:]

    def front_end_buf(self, buf, row0):
        u, r0 = self._rows(buf, row0)
        b, n = self.b, self.n
        Au = np.zeros(b * n, complex)
        Au[-n:] = self.P.up[b - 1] * u[b - r0]
        u[0 - r0:b - r0] = (self.TF - self.P.lu_HF.solve(Au)).reshape(b, n)

    def clone_context(self):
        import copy
        return copy.copy(self)                               # shares the factorisation, own parked front solution (self.TF)

    def front_tf_new(self):
        return [None]

    def front_tf_save(self, t):
        t[0] = self.TF.copy()

    def front_tf_load(self, t):
        self.TF = t[0]

    def sweep_forward_buf(self, buf, row0, m_from, m_to):
        u, r0 = self._rows(buf, row0)
        for m in range(m_from, m_to + 1):
            u[m - r0] -= self.P.lo[m] * self.P.T(m, u[m - 1 - r0])

    def sweep_backward_buf(self, buf, row0, m_from, m_to, diag):
        u, r0 = self._rows(buf, row0)
        n = self.n
        for m in range(m_from, m_to - 1, -1):
            v = u[m - 1 - r0].copy()
            if m < n:
                v += self.P.up[m - 1] * u[m - r0]
            u[m - 1 - r0] = u[m - 1 - r0] - self.P.T(m, v)

    # multi-vector entry points of the device backend: here one vector after the other (the schedule is what is tested)
    group = 2

    def batch_group(self, R):
        g = 1
        while g * 2 <= min(R, self.group):
            g *= 2
        return g

    def sweep_forward_multi_buf(self, bufs, row0, m_from, m_to):
        for bf in bufs:
            self.sweep_forward_buf(bf, row0, m_from, m_to)

    def sweep_backward_multi_buf(self, bufs, row0, m_from, m_to, diag):
        for bf in bufs:
            self.sweep_backward_buf(bf, row0, m_from, m_to, diag)

    def matvec_rows(self, j_lo, j_hi, x, south, north, out):
        n = self.n
        c1, c2, c3, c4, c5 = (c[j_lo:j_hi] for c in self.c)
        u = x.numpy().reshape(-1, n)
        y = c5 * u
        y[:, 1:] += c1[:, 1:] * u[:, :-1]
        y[:, :-1] += c2[:, :-1] * u[:, 1:]
        y[1:] += c3[1:] * u[:-1]
        y[:-1] += c4[:-1] * u[1:]
        if south is not None:
            y[0] += c3[0] * south.numpy()
        if north is not None:
            y[-1] += c4[-1] * north.numpy()
        out.copy_(torch.from_numpy(y.ravel()))


def _worker(rank, world, port, n, b, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helmholtz_preconditioner_b200.slab import SlabSolver
    omega = 2 * np.pi * 4 + 2j
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    be = NumpyBackend(n, b, omega, 61.0, c_mat)
    S = SlabSolver(be, n, b, rank, world)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    xl = torch.from_numpy(x[S.j0:S.j1].ravel().copy())
    out = torch.empty_like(xl)
    S.precond_apply(xl, out)
    ref = be.P.apply(x.ravel()).reshape(n, n)[S.j0:S.j1].ravel()
    e1 = np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref)
    S.matvec(xl, out)
    ref2 = orc.stencil_matvec(x.ravel(), **be.p).reshape(n, n)[S.j0:S.j1].ravel()
    e2 = np.linalg.norm(out.numpy() - ref2) / np.linalg.norm(ref2)
    # a batch of right-hand sides pipelined through the slabs: the same results, one by one
    xs = [rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) for _ in range(3)]
    pairs = [(torch.from_numpy(xx[S.j0:S.j1].ravel().copy()), torch.empty_like(xl)) for xx in xs]
    S.precond_apply_batch(pairs)
    e3 = 0.0
    for xx, (_, o) in zip(xs, pairs):
        refb = be.P.apply(xx.ravel()).reshape(n, n)[S.j0:S.j1].ravel()
        e3 = max(e3, np.linalg.norm(o.numpy() - refb) / np.linalg.norm(refb))
    ret[rank] = (e1, e2, S.m_lo, S.m_hi, e3)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_schedule_matches_oracle(world):
    n, b = 40, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, b, ret), nprocs=world, join=True)
    strips = []
    for r in range(world):
        e1, e2, m_lo, m_hi, e3 = ret[r]
        assert e1 < 1e-12 and e2 < 1e-13 and e3 < 1e-12, (r, e1, e2, e3)
        strips += list(range(m_lo, m_hi + 1))
    assert strips == list(range(b + 1, n + 1))      # every strip has exactly one owner


def test_slab_bounds():
    from helmholtz_preconditioner_b200.slab import slab_bounds
    assert slab_bounds(4096, 12, 8) == [512 * r for r in range(9)]
    with pytest.raises(ValueError):
        slab_bounds(40, 12, 8)


# ---- the Krylov host logic over gloo: typed requests served in batches, one all-reduce per Arnoldi column ------------
class NumpyVectors:
    """DeviceVectors (helmholtz_preconditioner_b200/gmres.py) on CPU tensors: same interface, reductions over gloo."""

    def __init__(self, group=None):
        self.allreduces = 0
        self.group = group

    def reserve(self, restart):
        pass

    def _sum(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.complex128)))
        dist.all_reduce(torch.view_as_real(t), group=self.group)
        self.allreduces += 1
        return t.numpy()

    def norm(self, x):
        return self.norm_batch([x])[0]

    def norm_batch(self, xs):
        v = self._sum([np.vdot(x.numpy(), x.numpy()) for x in xs])
        return [float(np.sqrt(t.real)) for t in v]

    def scale_copy(self, a, x, y):
        y.copy_(torch.from_numpy(a * x.numpy()))

    def axpy(self, a, x, y):
        y.numpy()[:] += a * x.numpy()

    def mgs(self, V, k, w):
        return self.mgs_batch([(V, k, w)])[0]

    def mgs_batch(self, items):
        R, k = len(items), items[0][1]
        h0 = self._sum([np.vdot(w.numpy(), w.numpy()) for _, _, w in items])
        H = np.zeros((k, R), complex)
        for j in range(k):
            d = self._sum([np.vdot(V[j].numpy(), w.numpy()) for V, _, w in items])
            H[j] = d
            for i, (V, _, w) in enumerate(items):
                w.numpy()[:] -= d[i] * V[j].numpy()
        h1 = self._sum([np.vdot(w.numpy(), w.numpy()) for _, _, w in items])
        return [(H[:, i].copy(), float(np.sqrt(h1[i].real)), float(np.sqrt(h0[i].real))) for i in range(R)]

    def combine(self, V, y, x):
        x.numpy()[:] += np.asarray(y) @ V[:len(y)].numpy()


def _gmres_worker(rank, world, port, n, b, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helmholtz_preconditioner_b200.slab import SlabSolver
    from helmholtz_preconditioner_b200.gmres import gmres, gmres_batch
    omega = 2 * np.pi * 4 + 2j
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    be = NumpyBackend(n, b, omega, 61.0, c_mat)
    S = SlabSolver(be, n, b, rank, world)
    fs = [f_mat, np.roll(f_mat, 7, axis=1), np.roll(f_mat, -5, axis=1)]
    loc = [torch.from_numpy(np.ascontiguousarray(f[S.j0:S.j1].ravel().astype(np.complex128))) for f in fs]
    vec = NumpyVectors()
    kw = dict(vec=vec, rtol=1e-3, restart=20, maxiter=12, nglobal=n * n)
    batch = gmres_batch(lambda x, o: S.matvec(x, o), lambda reqs: S.precond_apply_batch(reqs), loc,
                        matvec_batch=lambda reqs: S.matvec_batch(reqs), **kw)
    n_batch = vec.allreduces
    vec.allreduces = 0
    one = [gmres(lambda x, o: S.matvec(x, o), lambda x, o: S.precond_apply(x, o), f, **kw) for f in loc]
    n_one = vec.allreduces
    ret[rank] = dict(j0=S.j0, j1=S.j1, batch=[(u.numpy().copy(), info, hist) for u, info, hist in batch],
                     one=[(u.numpy().copy(), info, hist) for u, info, hist in one], n_batch=n_batch, n_one=n_one)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_distributed_gmres_batch_matches_oracle(world):
    """gmres_batch over slabs (gloo): every right-hand side follows the oracle's scipy-restated GMRES, the batched
    drivers give what the one-by-one driver gives, with 1/R of the all-reduces."""
    n, b = 40, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gmres_worker, args=(world, _free_port(), n, b, ret), nprocs=world, join=True)
    omega = 2 * np.pi * 4 + 2j
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    h = 1 / (n + 1)
    A = orc.build_A_matrix(b, 61.0, b * h, omega, h, n, c_mat)
    P = orc.SweepingPreconditioner(b, 61.0, b * h, omega, h, n, c_mat)          # reference diagonal, as the stand-in
    fs = [f_mat, np.roll(f_mat, 7, axis=1), np.roll(f_mat, -5, axis=1)]
    for i, f in enumerate(fs):
        u0, info0, hist0 = orc.gmres_scipy_restated(lambda x: A @ x, P.apply, f.ravel().astype(np.complex128), rtol=1e-3, maxiter=12)
        for kind in ("batch", "one"):
            u = np.concatenate([ret[r][kind][i][0] for r in range(world)])
            info, hist = ret[0][kind][i][1], ret[0][kind][i][2]
            assert info == info0 and len(hist) == len(hist0)
            assert np.allclose(hist, hist0, rtol=1e-8)
            assert np.linalg.norm(u - u0) / np.linalg.norm(u0) < 1e-8
    assert ret[0]["n_batch"] * 3 == ret[0]["n_one"]


# ---- groups of right-hand sides as independent pipelines (slab.GroupPipeline): threads, one process group per group ---
def _group_worker(rank, world, port, n, b, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helmholtz_preconditioner_b200.slab import SlabSolver, GroupPipeline
    from helmholtz_preconditioner_b200.gmres import gmres_batch
    omega = 2 * np.pi * 4 + 2j
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    be = NumpyBackend(n, b, omega, 61.0, c_mat)
    S = SlabSolver(be, n, b, rank, world)
    fs = [np.roll(f_mat, s, axis=1) for s in (0, 7, -5, 3, -9, 11)]
    loc = [torch.from_numpy(np.ascontiguousarray(f[S.j0:S.j1].ravel().astype(np.complex128))) for f in fs]
    kw = dict(rtol=1e-3, restart=20, maxiter=9, nglobal=n * n)
    pipe = GroupPipeline(S, 3, device="cpu", backend="gloo")
    outs = [[torch.zeros_like(x) for x in grp] for grp in (loc[0:2], loc[2:4], loc[4:6])]
    res = pipe.gmres([loc[0:2], loc[2:4], loc[4:6]], lambda nloc, pg: NumpyVectors(pg), host_out=outs, **kw)
    pipe.close()
    for grp, og in zip(res, outs):                           # the solutions also arrive in the buffers handed in
        for (u, _, _), o in zip(grp, og):
            assert torch.equal(u, o)
    vec = NumpyVectors()
    lock = gmres_batch(lambda x, o: S.matvec(x, o), lambda reqs: S.precond_apply_batch(reqs), loc, vec=vec,
                       matvec_batch=lambda reqs: S.matvec_batch(reqs), **kw)
    flat = [r for grp in res for r in grp]
    ret[rank] = dict(pipe=[(u.numpy().copy(), info, hist) for u, info, hist in flat],
                     lock=[(u.numpy().copy(), info, hist) for u, info, hist in lock])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_group_pipeline_matches_lock_step(world):
    """independent groups (threads, own process groups) give what the lock-step batch gives (to rounding: the all-reduce of
    another process group may sum the ranks in another order)"""
    n, b = 40, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_group_worker, args=(world, _free_port(), n, b, ret), nprocs=world, join=True)
    for r in range(world):
        for (u, info, hist), (u0, info0, hist0) in zip(ret[r]["pipe"], ret[r]["lock"]):
            assert info == info0 and len(hist) == len(hist0) and np.allclose(hist, hist0, rtol=1e-10)
            assert np.linalg.norm(u - u0) / np.linalg.norm(u0) < 1e-10


def test_pipeline_model_sanity():
    """tools/pipeline_model.py: one rank and one group is the plain sum of the parts; the asynchronous schedule is never
    slower than lock step in the model; more groups per rank raise the utilisation"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pipeline_model", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "pipeline_model.py"))
    pm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm)
    one = pm.simulate(1, 1, "lockstep", handover_ms=0.0, sync_ms=0.0)
    assert abs(one - (2 * 31.7 * 1.1 + 19.2)) < 0.5
    for N in (2, 4, 8):
        a, l = pm.simulate(N, N, "async"), pm.simulate(N, N, "lockstep")
        assert a <= l * 1.02
        assert pm.simulate(N, 2 * N, "async") / 2 < a
