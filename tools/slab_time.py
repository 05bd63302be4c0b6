"""torchrun --nproc-per-node N tools/slab_time.py : where does a batched, pipelined preconditioner application spend
its time (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import helmholtz_preconditioner_b200 as hp
from helmholtz_preconditioner_b200.slab import distributed_gmres_setup

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n, b = 4096, 12
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * world
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
S = distributed_gmres_setup(n, b, omega, 100.0, c_mat, rank, world, None, dev)
xs = [torch.randn(S.rows * n, dtype=torch.complex128, device=dev) for _ in range(R)]
outs = [torch.empty_like(x) for x in xs]
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.time()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.time() - t0) / reps * 1e3
t_one = timed(lambda: S.precond_apply(xs[0], outs[0]))
t_batch = timed(lambda: S.precond_apply_batch(list(zip(xs, outs))))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m_to = min(S.m_hi, n - 1)
e0.record(); S.s.sweep_forward_buf(S.buf, S.j0 - 1, S.m_lo, m_to); e1.record(); torch.cuda.synchronize()
t_f = e0.elapsed_time(e1)
e0.record(); S.s.sweep_backward_buf(S.buf, S.j0 - 1, S.m_hi, S.m_lo, "reference"); e1.record(); torch.cuda.synchronize()
t_b = e0.elapsed_time(e1)
print(f"rank {rank}: strips {S.m_lo}..{S.m_hi}  forward sweep {t_f:.2f} ms  backward {t_b:.2f} ms | one rhs {t_one:.1f} ms | batch of {R}: {t_batch:.1f} ms "
      f"= {t_batch / R:.1f} ms per rhs (ideal {(world - 1 + R) * (t_f + t_b) / R:.1f})", flush=True)
from helmholtz_preconditioner_b200.gmres import DeviceVectors
vec = DeviceVectors(xs[0].numel(), dev, group=dist.group.WORLD)
V = torch.randn(8, xs[0].numel(), dtype=torch.complex128, device=dev)
t_mv = timed(lambda: [S.matvec(x, o) for x, o in zip(xs, outs)], 3)
t_mgs = timed(lambda: [vec.mgs(V, 4, o) for o in outs], 3)
t_nrm = timed(lambda: [vec.norm(o) for o in outs], 3)
t_sc = timed(lambda: [vec.scale_copy(0.5, x, o) for x, o in zip(xs, outs)], 3)
t_alloc = timed(lambda: [torch.empty((21, xs[0].numel()), dtype=torch.complex128, device=dev) for _ in range(R)], 3)
if rank == 0:
    print(f"per batch of {R}: matvec {t_mv:.2f} ms, mgs(k=4) {t_mgs:.2f} ms, norm {t_nrm:.2f} ms, scale_copy {t_sc:.2f} ms, basis alloc {t_alloc:.2f} ms", flush=True)
dist.destroy_process_group()
