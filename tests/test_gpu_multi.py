"""Multi-vector sweep kernel (csrc/hp_sweep4m.cu): algo2_4 applied to R right-hand sides in one pass over the strip
generators must give, per right-hand side, what the single-vector kernel gives (and therefore what the oracle gives:
tests/test_gpu_parity.py, tests/test_gpu_large.py).  Runs on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import helmholtz_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def hp():
    import helmholtz_preconditioner_b200 as hp
    hp.load()
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    return hp


def rel(a, b):
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()


@pytest.mark.parametrize("n,b,P,K", [(300, 12, 6, 4), (255, 12, 0, 0), (1024, 12, 0, 0), (130, 20, 6, 5), (200, 5, 7, 2)])
def test_multi_matches_single(hp, n, b, P, K):
    omega = 2 * np.pi * (n / 10) + 2j
    c_mat = orc.init_c1_f1(omega, n)[0]
    s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(P=P, K=K, layout="cluster")
    mm = s.multi_max
    assert mm in (1, 2, 4, 8)
    if mm == 1:
        L = s.layout()
        assert L["CW"] > 32 or L["P"] - 1 > 32, "a partition with CW <= 32 and P-1 <= 32 must support the multi-vector kernel"
        pytest.skip("partition does not fit the multi-vector kernel")
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = [torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g) for _ in range(8)]
    for diag in ("reference", "paper"):
        singles = [s.precond_apply(x, diag=diag) for x in xs]
        for R in (1, 2, 4, 8):
            if R > mm:
                continue
            outs = [torch.empty_like(x) for x in xs[:R]]
            s.precond_apply_multi(xs[:R], outs, diag=diag)
            for o, ref in zip(outs, singles):
                assert rel(o, ref) < 1e-13, (diag, R)
        # in place, and a batch that is not a power of two (8 -> 4 + 1 ...)
        pairs = [(x, torch.empty_like(x)) for x in xs[:5]]
        s.precond_apply_batch(pairs, diag=diag)
        for (_, o), ref in zip(pairs, singles):
            assert rel(o, ref) < 1e-13
    s.check_status()
    s.close()


def test_multi_vs_oracle(hp):
    """directly against the oracle (SuperLU), 4 right-hand sides at once"""
    n, b = 150, 12
    omega = 2 * np.pi * 15 + 2j
    c_mat, f_mat = orc.init_c2_f2(omega, n)
    h = 1 / (n + 1)
    s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(layout="cluster")
    if s.multi_max < 4:
        pytest.skip("partition does not fit the multi-vector kernel")
    rng = np.random.default_rng(2)
    xs = [f_mat.ravel().astype(np.complex128)] + [rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n) for _ in range(3)]
    for diag in ("reference", "paper"):
        P = orc.SweepingPreconditioner(b, 60.0, b * h, omega, h, n, c_mat, diag=diag)
        dx = [torch.from_numpy(x).cuda() for x in xs]
        outs = [torch.empty_like(x) for x in dx]
        s.precond_apply_multi(dx, outs, diag=diag)
        for x, o in zip(xs, outs):
            ref = P.apply(x)
            assert np.linalg.norm(o.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-12
    s.check_status()
    s.close()


@pytest.mark.parametrize("front", ["blockdiag", "coupled"])
def test_group_pipeline_one_gpu(hp, front):
    """slab.GroupPipeline on one GPU (world 1): three groups of right-hand sides, each with its own thread, stream and solver
    context (hp_context_clone), their sweeps interleaving on the device, must give bit for bit what the lock-step batch
    on the parent solver gives, and the contexts must not disturb each other (exchange ring, parked front solutions)."""
    from helmholtz_preconditioner_b200.slab import SlabSolver, GroupPipeline
    from helmholtz_preconditioner_b200.gmres import DeviceVectors, gmres_batch
    n, b = 1024, 12
    omega = 2 * np.pi * (n / 10) + 2j
    c_mat, f_mat = hp.init_layered_f1(omega, n)
    s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner(front=front)
    S = SlabSolver(s, n, b, 0, 1, None, device=s.device)
    sizes = [8, 8, 3]
    fs = [torch.from_numpy(np.roll(f_mat, 17 * i, axis=1).ravel().astype(np.complex128)).cuda() for i in range(sum(sizes))]
    groups, i = [], 0
    for g in sizes:
        groups.append(fs[i:i + g]); i += g
    kw = dict(rtol=1e-3, restart=20, maxiter=7, nglobal=n * n)
    pipe = GroupPipeline(S, len(sizes))
    res = pipe.gmres(groups, lambda nloc, pg: DeviceVectors(nloc, s.device, group=pg), diag="paper", **kw)
    torch.cuda.synchronize()
    assert pipe.sweep_status() == 0
    pipe.close()
    for grp, rg in zip(groups, res):
        vec = DeviceVectors(n * n, s.device)
        lock = gmres_batch(lambda x, o: S.matvec(x, o), lambda reqs: S.precond_apply_batch(reqs, diag="paper"), grp, vec=vec,
                           matvec_batch=lambda reqs: S.matvec_batch(reqs), **kw)
        for (u, info, hist), (u0, info0, hist0) in zip(rg, lock):
            assert info == info0 and hist == hist0
            assert torch.equal(u, u0)
    s.check_status()
    s.close()


@pytest.mark.parametrize("n,k,R", [(1000, 1, 1), (70001, 4, 3), (300000, 5, 8), (2097152, 13, 8), (123457, 20, 2), (5000, 0, 2)])
def test_cgs_pass_kernels(hp, n, k, R):
    """csrc/hp_cgs.cu: dots of w against k basis vectors + |w|^2, the block update, for R systems in one launch, against torch"""
    import ctypes as C
    from helmholtz_preconditioner_b200 import _lib
    lib = _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(n + k)
    rnd = lambda *s: torch.randn(*s, dtype=torch.complex128, device="cuda", generator=g)      # noqa: E731
    Vs = [rnd(21, n) for _ in range(R)]
    ws = [rnd(n) for _ in range(R)]
    w0 = [w.clone() for w in ws]
    out = torch.zeros((2, R, 21), dtype=torch.complex128, device="cuda")
    arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])                          # noqa: E731
    st = torch.cuda.current_stream().cuda_stream
    # pass A: dots and norm, w untouched
    _lib.check(lib.hp_cgs_pass(R, n, k, arr(Vs), Vs[0].stride(0), arr(ws), None, arr([out[0, i] for i in range(R)]), 0, 1, st), "hp_cgs_pass")
    for i in range(R):
        assert torch.equal(ws[i], w0[i])
        d = Vs[i][:k].conj() @ w0[i]
        if k:
            assert (torch.linalg.norm(out[0, i, :k] - d) / torch.linalg.norm(Vs[i][:k].abs() @ w0[i].abs())).item() < 1e-13
        assert abs(out[0, i, k].real.item() - torch.linalg.norm(w0[i]).item() ** 2) < 1e-12 * out[0, i, k].real.item()
    if k == 0:
        return
    # pass B: update with the coefficients of pass A, dots and norm of the updated vector; pass C: update only + norm
    for upd_dots, slot in ((1, 1), (0, 1)):
        before = [w.clone() for w in ws]
        coef = out[0].clone() if upd_dots else out[1].clone()
        _lib.check(lib.hp_cgs_pass(R, n, k, arr(Vs), Vs[0].stride(0), arr(ws), arr([coef[i] for i in range(R)]),
                                   arr([out[slot, i] for i in range(R)]), 1, upd_dots, st), "hp_cgs_pass")
        for i in range(R):
            ref = before[i] - coef[i, :k] @ Vs[i][:k]
            assert (torch.linalg.norm(ws[i] - ref) / torch.linalg.norm(ref)).item() < 1e-13
            if upd_dots:
                d = Vs[i][:k].conj() @ ref
                assert (torch.linalg.norm(out[slot, i, :k] - d) / torch.linalg.norm(Vs[i][:k].abs() @ ref.abs())).item() < 1e-13      # scale of the sums
            assert abs(out[slot, i, k].real.item() - torch.linalg.norm(ref).item() ** 2) < 1e-12 * out[slot, i, k].real.item()


def test_cgs2_matches_mgs(hp):
    """DeviceVectors on 'distributed' vectors (a process group of one rank): classical Gram-Schmidt twice against the
    modified Gram-Schmidt path: same coefficients and the same orthogonalised vector to rounding"""
    import os
    import socket
    import torch.distributed as dist
    from helmholtz_preconditioner_b200.gmres import DeviceVectors
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        n, k, R = 200003, 9, 5
        g = torch.Generator(device="cuda").manual_seed(3)
        items = []
        for _ in range(R):
            Q, _r = torch.linalg.qr(torch.randn(n, k, dtype=torch.complex128, device="cuda", generator=g))
            V = torch.zeros(21, n, dtype=torch.complex128, device="cuda")
            V[:k] = Q.T
            items.append((V, k, torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)))
        a = DeviceVectors(n, "cuda", group=dist.group.WORLD, orth="mgs")
        b = DeviceVectors(n, "cuda", group=dist.group.WORLD, orth="cgs2")
        ia = [(V, k, w.clone()) for V, k, w in items]
        ib = [(V, k, w.clone()) for V, k, w in items]
        ra, rb = a.mgs_batch(ia), b.mgs_batch(ib)
        for (ha, na1, na0), (hb, nb1, nb0), (_, _, wa), (_, _, wb) in zip(ra, rb, ia, ib):
            assert np.linalg.norm(ha - hb) / np.linalg.norm(ha) < 1e-13
            assert abs(na1 - nb1) < 1e-13 * na1 and abs(na0 - nb0) < 1e-13 * na0
            assert (torch.linalg.norm(wa - wb) / torch.linalg.norm(wa)).item() < 1e-13
    finally:
        dist.destroy_process_group()
