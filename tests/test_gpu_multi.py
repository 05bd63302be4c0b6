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
