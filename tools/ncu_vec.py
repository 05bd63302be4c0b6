"""ncu target: the bandwidth-bound kernels of the path at 4096^2, a few launches each on inputs larger than L2
(stencil SpMV, CSR assembly, fused Gram-Schmidt steps, basis combination, reductions).  No strips are factored.
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:'hp_(stencil|assemble|axpy|combine|reduce|scale)' python tools/ncu_vec.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
from helmholtz_preconditioner_b200.gmres import DeviceVectors

n, b = 4096, 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
N = n * n
g = torch.Generator(device="cuda").manual_seed(3)
V = torch.randn(21, N, dtype=torch.complex128, device="cuda", generator=g)
w = torch.randn(N, dtype=torch.complex128, device="cuda", generator=g)
y = torch.empty_like(w)
vec = DeviceVectors(N, w.device)
for i in range(3):
    s.matvec(V[i], y)
A = s.assemble_csr()
A = s.assemble_csr()
vec.mgs(V, 20, w)
vec.combine(V, np.ones(20, complex) / 20, y)
vec.norm(w)
vec.scale_copy(0.5, w, y)
torch.cuda.synchronize()
print("done")
