"""Developer probe: device time of the pieces of one preconditioned Krylov iteration.  (GPU box)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
from helmholtz_preconditioner_b200.gmres import DeviceVectors, gmres

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = int(sys.argv[2]) if len(sys.argv) > 2 else 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner()
N = n * n
f = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
u = f.clone()
V = torch.randn(21, N, dtype=torch.complex128, device="cuda")
w = torch.randn(N, dtype=torch.complex128, device="cuda")
vec = DeviceVectors(N, f.device)


def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1) / reps:9.3f} ms   wall {(time.perf_counter() - t0) / reps * 1e3:9.3f} ms")


timed("front_begin", lambda: s.front_begin(u))
timed("sweep_forward", lambda: s.sweep_forward(u, b + 1, n - 1))
timed("sweep_backward", lambda: s.sweep_backward(u, n, b + 1))
timed("front_end", lambda: s.front_end(u))
timed("precond_apply", lambda: s.precond_apply(f, out=u))
timed("matvec", lambda: s.matvec(f, u), 10)
timed("mgs k=10", lambda: vec.mgs(V, 10, w))
timed("mgs k=20", lambda: vec.mgs(V, 20, w))
timed("norm", lambda: vec.norm(w), 10)
timed("scale_copy", lambda: vec.scale_copy(0.5, w, u), 10)
timed("combine k=20", lambda: vec.combine(V, np.ones(20, complex), u))
mv = lambda x, out: s.matvec(x, out)
ps = lambda x, out: s.precond_apply(x, out=out)
timed("gmres 20 iterations", lambda: gmres(mv, ps, f, vec=vec, rtol=0.0, restart=20, maxiter=20), 1)
print("status", s.sweep_status())
