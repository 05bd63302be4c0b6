"""Developer tool: strip setup time at n^2 under the developer switches of csrc/hp_setup.cu, several setups per
process (the first one of a process runs on a cold GPU and is reported separately).  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
SW = ["HP_CHAIN_SMEM", "HP_CHAIN_UNROLL", "HP_SETUP_THREAD", "HP_LEAF_NOPIPE", "HP_LEAF_CTA", "HP_CORNER_WARP", "HP_SEP_ROWS_BUF"]
QUICK = len(sys.argv) > 2 and sys.argv[2] == "quick"
configs = [("default", []), ("old chain", ["HP_CHAIN_SMEM"]), ("unrolled inverse", ["HP_CHAIN_UNROLL"]),
           ("thread sep/corner", ["HP_SETUP_THREAD"]), ("leaf CTA-paced", ["HP_LEAF_CTA"]),
           ("leaf CTA, not piped", ["HP_LEAF_CTA", "HP_LEAF_NOPIPE"]),
           ("all old", ["HP_CHAIN_SMEM", "HP_SETUP_THREAD", "HP_LEAF_CTA", "HP_LEAF_NOPIPE"]), ("default", [])]
if QUICK:
    configs = [("default", []), ("corner: leaf per warp", ["HP_CORNER_WARP"]), ("N rows via buffer", ["HP_SEP_ROWS_BUF"]), ("default", [])]
ref = None
x = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
for name, sw in configs:
    for k in SW:
        os.environ.pop(k, None)
    for k in sw:
        os.environ[k] = "1"
    ts = []
    for rep in range(2):
        s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
        s.setup_preconditioner()
        ts.append(s.setup_ms)
        if rep == 1:
            y = s.precond_apply(x).clone()
            if ref is None:
                ref = y
            d = float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref))
        s.close()
        del s
    print(f"{name:20s} setup ms {ts[0]:8.1f} {ts[1]:8.1f}   |M f - M f(default)|/|M f| = {d:.2e}", flush=True)
