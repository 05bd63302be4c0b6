"""torchrun --nproc-per-node 2 tools/slab_check.py [n] : slab-decomposed operator and preconditioner on 2+ GPUs
against the single-GPU result computed on rank 0 (GPU box, needs >= 2 GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import helmholtz_preconditioner_b200 as hp
from helmholtz_preconditioner_b200.slab import distributed_gmres_setup
from helmholtz_preconditioner_b200.gmres import DeviceVectors, gmres

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
S = distributed_gmres_setup(n, b, omega, 100.0, c_mat, rank, world, None, dev)
rng = np.random.default_rng(11)
x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
xl = torch.from_numpy(x[S.j0:S.j1].ravel().copy()).to(dev)
out = torch.empty_like(xl)
res = {}
S.precond_apply(xl, out); res["M"] = out.clone()
S.matvec(xl, out); res["A"] = out.clone()
# batch of right-hand sides pipelined through the slabs == one by one
xs = [torch.from_numpy((rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))[S.j0:S.j1].ravel().copy()).to(dev) for _ in range(5)]
singles = []
for xx in xs:
    o = torch.empty_like(xx); S.precond_apply(xx, o); singles.append(o)
pairs = [(xx, torch.empty_like(xx)) for xx in xs]
S.precond_apply_batch(pairs)
eb = max((torch.linalg.norm(o - s1_) / torch.linalg.norm(s1_)).item() for (_, o), s1_ in zip(pairs, singles))
ebt = torch.tensor([eb], device=dev); dist.all_reduce(ebt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"batch of 5 pipelined vs one by one: max rel diff {ebt.item():.2e}")
assert ebt.item() < 1e-13
vec = DeviceVectors(xl.numel(), dev, group=dist.group.WORLD)
fl = torch.from_numpy(f_mat[S.j0:S.j1].ravel().astype(np.complex128)).to(dev)
u, info, hist = gmres(lambda a, o: S.matvec(a, o), lambda a, o: S.precond_apply(a, o, diag="paper"), fl, vec=vec,
                      rtol=1e-3, restart=20, maxiter=15, nglobal=n * n)
res["u"] = u
gath = {k: [torch.empty((S.R[r + 1] - S.R[r]) * n, dtype=torch.complex128, device=dev) for r in range(world)] for k in res}
for k in res:
    dist.all_gather(gath[k], res[k]) if len({g.numel() for g in gath[k]}) == 1 else None
if rank == 0:
    s1 = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat, device=dev).setup_preconditioner()
    xf = torch.from_numpy(x.ravel()).to(dev)
    rel = lambda a, c: (torch.linalg.norm(a - c) / torch.linalg.norm(c)).item()
    eM = rel(torch.cat(gath["M"]), s1.precond_apply(xf))
    eA = rel(torch.cat(gath["A"]), s1.matvec(xf))
    v1 = DeviceVectors(n * n, dev)
    f1 = torch.from_numpy(f_mat.ravel().astype(np.complex128)).to(dev)
    u1, info1, hist1 = gmres(lambda a, o: s1.matvec(a, o), lambda a, o: s1.precond_apply(a, out=o, diag="paper"), f1, vec=v1,
                             rtol=1e-3, restart=20, maxiter=15)
    eU = rel(torch.cat(gath["u"]), u1)
    print(f"slab check n={n} world={world}: |M|err {eM:.2e} |A|err {eA:.2e} gmres iters {len(hist)} vs {len(hist1)} |u|err {eU:.2e} "
          f"hist rel {max(abs(a - c) / c for a, c in zip(hist, hist1)):.2e}")
    assert eM < 1e-11 and eA < 1e-13 and len(hist) == len(hist1) and eU < 1e-8
    print("SLAB CHECK OK")
dist.destroy_process_group()
