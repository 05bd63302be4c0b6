"""Developer probe: globaltimer stamps of the critical cycle of the pipelined sweep (strips 512..575).  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
n, b = 4096, 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
s.setup_preconditioner(m_lo=b + 1, m_hi=b + 800)
L = s.layout(); G, K = L["G"], L["K"]
u = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
s.sweep_forward(u, b + 1, b + 799); torch.cuda.synchronize()
s.lib.hp_debug_phases(s.handle, 1, None)
s.sweep_forward(u, b + 1, b + 799); torch.cuda.synchronize()
out = np.zeros(G * (16 + 1024), dtype=np.int64)
s.lib.hp_debug_phases(s.handle, 0, out.ctypes.data)
st = out[G * 16:].reshape(G, 64, 4).astype(np.float64)
red = np.arange(G) % K == 0
rows = np.arange(G) * L["NR"] < L["NS"]
t0 = st[red][:, :, 0]     # reducer got XS(t-1)+GPb(t)
t1 = st[red][:, :, 1]     # reducer published GR(t)
t2 = st[rows][:, :, 2]    # CTA got all GR(t)
t3 = st[rows][:, :, 3]    # CTA published XS(t) rows
T = slice(5, 60)
per = np.diff(t3.max(0))[T].mean()
print("period ns", per)
print("XS(t-1) last published -> reducers have inputs: first %.0f  mean %.0f  last %.0f ns" % tuple(f((t0[:, 1:] - t3.max(0)[None, :-1])[:, T]) for f in (lambda x: x.min(0).mean(), lambda x: x.mean(), lambda x: x.max(0).mean())))
print("reducer inputs -> GR published: mean %.0f max %.0f ns" % ((t1 - t0)[:, T].mean(), (t1 - t0)[:, T].max(0).mean()))
print("last GR published -> CTAs have all GR: first %.0f mean %.0f last %.0f ns" % tuple(f((t2 - t1.max(0)[None, :])[:, T]) for f in (lambda x: x.min(0).mean(), lambda x: x.mean(), lambda x: x.max(0).mean())))
print("GR received -> XS rows published: mean %.0f max %.0f ns" % ((t3 - t2)[:, T].mean(), (t3 - t2)[:, T].max(0).mean()))
print("spread of GR publish times across reducers: %.0f ns ; spread of XS publish across CTAs %.0f ns" % ((t1.max(0) - t1.min(0))[T].mean(), (t3.max(0) - t3.min(0))[T].mean()))
