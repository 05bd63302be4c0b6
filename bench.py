#!/usr/bin/env python
"""bench.py -- preconditioned Krylov iterations/s of the moving-PML sweeping-preconditioner Helmholtz solve.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is ONE preconditioned GMRES inner iteration of the reference's solve (code.py:516): one 5-point
stencil SpMV, one application of the sweeping preconditioner (algo2_4: front solves + forward and backward
sweeps over all n-b moving-PML strips), the modified Gram-Schmidt orthogonalisation and the host Givens
update.  K steps = restarted GMRES(20) run for exactly K inner iterations (rtol = 0), including the
solution update / true residual at every restart boundary.  The preconditioner is applied to the Krylov
vector (precond_input='vector'): per iteration this is exactly the work the reference does (it re-runs
algo2_4 on every call, code.py:510-511).

One JSON line is printed by rank 0.  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--grid-n", dest="n", type=int, default=0, help="interior grid points per side (default: 4096; --grid-n under torchrun, whose parser claims --n)")
    ap.add_argument("--b", type=int, default=12, help="PML width in grid points (reference: 12)")
    ap.add_argument("--ppw", type=float, default=10.0, help="grid points per wavelength")
    ap.add_argument("--const", type=float, default=100.0)
    ap.add_argument("--model", default="layered", choices=["layered", "constant", "c1"])
    ap.add_argument("--cpu-strips", type=int, default=6, help="strips timed for the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tts", action="store_true")
    ap.add_argument("--mp-mode", default="pipelined", choices=["pipelined", "weak"],
                    help="N > 1: 'pipelined' = the 4096^2 problem slab-decomposed, 8N right-hand sides sent through the slabs one "
                         "behind the other; 'weak' = one problem of 4096^2 points per GPU (n = 4096 sqrt(N)), one right-hand side")
    ap.add_argument("--rhs", type=int, default=0, help="right-hand sides in flight in the pipelined mode (default 8N)")
    return ap.parse_args()


def workload(args, world):
    """BASELINE.json configs: N=1 -> 'heterogeneous synthetic layered velocity model 4096^2, preconditioned
    solve, 1 B200' (the configuration the metric is quoted on); N>1 -> weak scaling, 4096^2 points per GPU."""
    n = args.n if args.n else (int(round(4096 * np.sqrt(world))) if args.mp_mode == "weak" else 4096)
    wave_num = n / args.ppw
    return dict(n=n, b=args.b, wave_num=wave_num, const=args.const, alpha=2.0, model=args.model)


def make_fields(w):
    import helmholtz_preconditioner_b200 as hp
    omega = 2 * np.pi * w["wave_num"] + 1j * w["alpha"]
    n = w["n"]
    if w["model"] == "layered":
        c_mat, f_mat = hp.init_layered_f1(omega, n)
    elif w["model"] == "constant":
        c_mat, f_mat = hp.init_const_f1(omega, n)
    else:
        c_mat, f_mat = hp.init_c1_f1(omega, n)
    return omega, c_mat, np.ascontiguousarray(f_mat.astype(np.complex128))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path, bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_iteration_sample(w, nstrips, c_mat, f_mat):
    """Time the reference algorithm (oracle port: scipy SuperLU strip solves as in code.py:345-385, CSR matvec)
    on `nstrips` of the n-b strips and extrapolate one preconditioned Krylov iteration linearly in the number
    of strips (every strip costs the same: 3 SuperLU solves of a bn x bn system, code.py:366-380)."""
    import scipy.sparse.linalg as spla
    from oracle import helmholtz_oracle as orc
    n, b = w["n"], w["b"]
    omega = 2 * np.pi * w["wave_num"] + 1j * w["alpha"]
    h = 1 / (n + 1)
    eta = b * h
    ms = np.linspace(b + 1, n, nstrips).astype(int)
    t_fac = t_sol = 0.0
    rng = np.random.default_rng(0)
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for m in ms:
        t0 = time.perf_counter()
        Hm = orc.get_Hm(int(m), b, w["const"], eta, omega, h, n, c_mat).tocsc()
        lu = spla.splu(Hm)
        t1 = time.perf_counter()
        t = np.zeros(b * n, dtype=np.complex128)
        for _ in range(3):                        # code.py:370, 375, 380
            t[-n:] = v
            lu.solve(t)[-n:]
        t2 = time.perf_counter()
        t_fac += t1 - t0
        t_sol += t2 - t1
    A = orc.build_A_matrix(b, w["const"], eta, omega, h, n, c_mat)
    x = f_mat.ravel()
    t0 = time.perf_counter()
    A @ x
    t_mv = time.perf_counter() - t0
    # orthogonalisation at the average Krylov index of GMRES(20): ~10 vdot + 10 axpy
    V = np.stack([x, x * 1j])
    t0 = time.perf_counter()
    for _ in range(5):
        hcoef = np.vdot(V[0], V[1])
        V[1] -= hcoef * V[0]
    t_orth = (time.perf_counter() - t0) * 2
    per_strip = t_sol / len(ms)
    t_iter = per_strip * (n - b) + t_mv + t_orth
    return dict(t_iter=t_iter, per_strip_solve=per_strip, per_strip_factor=t_fac / len(ms), t_matvec=t_mv,
                t_setup=t_fac / len(ms) * (n - b), strips=len(ms))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w = workload(args, world)
    omega, c_mat, f_mat = make_fields(w)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_iteration_sample(w, max(1, args.cpu_strips // 2), c_mat, f_mat)
        if i >= args.warmup:
            vals.append(r["t_iter"])
    t_iter = float(np.mean(vals))
    sample = (f"{max(1, args.cpu_strips // 2)} of {w['n'] - w['b']} strips per step (SuperLU factor + 3 solves each, "
              f"code.py:345-380) + 1 CSR matvec, extrapolated linearly to all strips")
    out = {"impl": "reference", "metric": "precond. Krylov iters/s at 4096^2 2D", "value": 1.0 / t_iter, "unit": "iters/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_iter,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
           "data": "synthetic", "config": config_dict(w, world),
           "cpu_baseline": {"value": 1.0 / t_iter, "unit": "iters/s", "cores": 1, "kind": "port", "sample": sample},
           "e2e": {"value": 1.0 / t_iter, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def config_dict(w, world):
    return {"workload": (f"2D heterogeneous synthetic layered velocity model {w['n']}^2, PML width {w['b']}, "
                         f"{w['n'] / w['wave_num']:.0f} points per wavelength, preconditioned GMRES(20) inner iterations"
                         if w["model"] == "layered" else f"2D {w['model']} velocity {w['n']}^2, PML width {w['b']}"),
            "n": w["n"], "b": w["b"], "wave_num": w["wave_num"], "const": w["const"], "alpha": w["alpha"],
            "restart": 20, "precond_input": "vector", "diag": "reference",
            "l2": "inputs larger than L2: the sweep streams the strip generators (GBs per step) once per step",
            "parallelism": f"slab{world}" if world > 1 else "single"}


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import helmholtz_preconditioner_b200 as hp
    from helmholtz_preconditioner_b200 import _lib
    from helmholtz_preconditioner_b200.gmres import DeviceVectors, gmres
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from helmholtz_preconditioner_b200 import slab
        return slab.bench_distributed(args, workload(args, world), make_fields, config_dict, ClockSampler)
    torch.cuda.set_device(local)
    lib = _lib.require_device()
    w = workload(args, 1)
    omega, c_mat, f_mat = make_fields(w)
    n, b = w["n"], w["b"]
    N = n * n
    # process start-up (CUDA context, loading the kernels of the library) is not part of the setup of a problem: a 96^2
    # solve brings both up before the clock starts
    om0 = 2 * np.pi * 9.6 + 2j
    c0, f0 = hp.init_layered_f1(om0, 96)
    s0 = hp.HelmholtzSolver(96, b, om0, w["const"], c0)
    s0.setup_preconditioner()
    s0.precond_apply(torch.from_numpy(f0.ravel().astype(np.complex128)).cuda())
    torch.cuda.synchronize()
    s0.close()
    del s0
    t0 = time.time()
    s = hp.HelmholtzSolver(n, b, omega, w["const"], c_mat)
    s.setup_preconditioner()
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    L = s.layout()
    f_host = torch.from_numpy(f_mat.ravel()).pin_memory()
    u_host = torch.empty(N, dtype=torch.complex128).pin_memory()
    f = f_host.cuda(non_blocking=True)
    vec = DeviceVectors(N, f.device)
    mv = lambda x, out: s.matvec(x, out)                      # noqa: E731
    ps = lambda x, out: s.precond_apply(x, out=out)           # noqa: E731

    def iterations(k, rhs):
        return gmres(mv, ps, rhs, vec=vec, rtol=0.0, atol=0.0, restart=20, maxiter=k)

    iterations(args.warmup, f)                                 # warm-up steps (untimed)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    # ---- device-resident timing: inputs already in HBM
    lib.hp_profile_enable(s.handle, 1)
    l0 = lib.hp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    u, info, hist = iterations(args.steps, f)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = lib.hp_launch_count() - l0
    import ctypes as C
    sw_ms, sw_n, sw_b = C.c_double(), C.c_int(), C.c_int64()
    lib.hp_profile_read(s.handle, C.byref(sw_ms), C.byref(sw_n), C.byref(sw_b))
    lib.hp_profile_enable(s.handle, 0)
    # ---- end to end through the public API with host buffers: pinned f -> device, K iterations, u -> host
    torch.cuda.synchronize()
    e0.record()
    f2 = f_host.cuda(non_blocking=True)
    u2, info2, hist2 = iterations(args.steps, f2)
    u_host.copy_(u2, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1) / 1e3
    clk = clocks.stop()

    # ---- secondary kernels of the path: matrix-free stencil SpMV and CSR assembly, device time per launch
    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record(); torch.cuda.synchronize()
        return a0.elapsed_time(a1) / 1e3 / reps
    xs = [torch.randn(N, dtype=torch.complex128, device=f.device) for _ in range(3)]   # 3 x 268 MB > L2
    ys = torch.empty_like(f)
    cnt = [0]
    def spmv():
        cnt[0] += 1
        s.matvec(xs[cnt[0] % 3], ys)
    t_spmv = timed(spmv, 30)
    spmv_bytes = N * (16 + 16 + 8)                      # x read, y written, kappa read
    nnz = 5 * N - 4 * n
    A_ = s.assemble_csr()
    t_asm = timed(lambda: lib.hp_assemble_csr(s.handle, A_.indptr.data_ptr(), A_.indices.data_ptr(), A_.data.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), 5)
    asm_bytes = nnz * (16 + 4) + (N + 1) * 4 + N * 8    # values + column indices + row pointers written, kappa read
    del A_, xs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = (sw_b.value / 1e9) / (sw_ms.value / 1e3) if sw_ms.value > 0 else 0.0
    traffic = None
    try:
        per_strip = json.load(open(os.path.join(ROOT, "profiles", "sweep_traffic.json")))["dram_bytes_per_strip"]
        # ncu figure per strip x the strips of an average timed launch (forward: n-b-1 strips, backward: n-b)
        traffic = per_strip * (n - b - 0.5) if (n, b) == (4096, 12) else None
    except Exception:
        pass
    out = {"metric": "precond. Krylov iters/s at 4096^2 2D", "value": args.steps / t_dev, "unit": "iters/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
           "data": "synthetic", "config": config_dict(w, 1),
           "e2e": {"value": args.steps / t_e2e, "unit": "iters/s", "h2d_bytes_per_step": f_host.numel() * 16 / args.steps,
                   "d2h_bytes_per_step": u_host.numel() * 16 / args.steps + 16 * 22,
                   "note": "host f -> device, K GMRES iterations, u -> host; per-iteration Hessenberg columns come back every step"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "hp_sweep4_kernel" if L.get("colN") else "hp_sweep2_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "launches_timed": sw_n.value, "avg_launch_ms": sw_ms.value / max(sw_n.value, 1),
                        "algorithmic_bytes_per_launch": sw_b.value / max(sw_n.value, 1),
                        "share_of_step": (sw_ms.value / 1e3) / t_dev},
           "spmv": {"kernel": "hp_stencil_matvec_kernel", "ms": 1e3 * t_spmv, "GB/s": spmv_bytes / t_spmv / 1e9,
                    "frac_of_hbm_peak": spmv_bytes / t_spmv / 1e9 / peak, "bytes_per_point": 40},
           "assembly": {"kernel": "hp_assemble_csr_kernel", "ms": 1e3 * t_asm, "GB/s": asm_bytes / t_asm / 1e9,
                        "frac_of_hbm_peak": asm_bytes / t_asm / 1e9 / peak, "nnz": nnz},
           "clocks": clk,
           "setup": {"seconds_wall": t_setup, "strip_factor_ms_device": s.setup_ms, "factor_bytes": s.precond_bytes,
                     "note": "wall time of HelmholtzSolver(...) + setup_preconditioner() in a process whose CUDA context and kernels are already loaded",
                     "partition": {k: int(L[k]) for k in ("P", "K", "G", "QP", "CW", "NS", "NR", "PK", "colN", "NRQ", "NXG")}},
           "residual_last": hist[-1] if hist else None}
    if not args.no_tts:
        # time to solution of the reference's literal call (code.py:510-516: M ignores its argument)
        torch.cuda.synchronize()
        t0 = time.time()
        r = hp.run_solver(n, b, w["wave_num"], w["const"], w["alpha"], c_mat=c_mat, f_mat=f_mat, solver=s, verbose=False)
        torch.cuda.synchronize()
        out["time_to_solution"] = {"setup_s": t_setup, "solve_s": time.time() - t0, "niter": r.niter, "info": r.info,
                                   "mode": "reference literal (precond_input='rhs', rtol=1e-3)"}
    if not args.no_cpu:
        c = cpu_iteration_sample(w, args.cpu_strips, c_mat, f_mat)
        out["cpu_baseline"] = {"value": 1.0 / c["t_iter"], "unit": "iters/s", "cores": 1, "kind": "port",
                               "sample": (f"{c['strips']} of {n - b} strips (SuperLU factor + 3 solves each, as code.py:345-380) "
                                          f"+ 1 CSR matvec on the host, extrapolated linearly to all strips"),
                               "per_strip_solve_s": c["per_strip_solve"], "per_strip_factor_s": c["per_strip_factor"],
                               "setup_s_extrapolated": c["t_setup"]}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
