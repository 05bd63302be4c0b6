"""Developer probe: how many GMRES(20) iterations the well-defined variants need to reach rtol 1e-3.  (GPU box)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

out = []
cases = [(1023, 12, 128, 100.0, "c1f1", 400), (4096, 12, 409.6, 100.0, "layered", 600)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1:]]
for n, b, wn, const, model, cap in cases:
    omega = 2 * np.pi * wn + 2j
    c_mat, f_mat = (hp.init_c1_f1(omega, n) if model == "c1f1" else hp.init_layered_f1(omega, n))
    s = hp.HelmholtzSolver(n, b, omega, const, c_mat).setup_preconditioner()
    A = s.assemble_csr()
    f = torch.from_numpy(np.ascontiguousarray(f_mat.ravel().astype(np.complex128))).cuda()
    for diag, mx in (("paper", cap), ("reference", 100)):
        t0 = time.time()
        r = hp.run_solver(n, b, wn, const, 2, c_mat=c_mat, f_mat=f_mat, solver=s, diag=diag, precond_input="vector",
                          maxiter=mx, verbose=False)
        torch.cuda.synchronize()
        dt = time.time() - t0
        res = (f - A.matvec(r.u)).norm().item() / f.norm().item()
        rec = dict(n=n, model=model, diag=diag, niter=r.niter, info=r.info, seconds=dt, true_res=res,
                   hist_first=r.residuals[:3], hist_last=r.residuals[-3:], status=s.sweep_status())
        print(json.dumps(rec), flush=True)
        out.append(rec)
    s.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/conv_probe.json", "w"), indent=1)
