"""Developer tool: per-kernel DRAM traffic / duration summary of the ncu metrics pass over tools/ncu_vec.py
(bandwidth-bound kernels of the path at 4096^2), with the algorithmic bytes of each kernel next to the measured traffic.
usage: python tools/ncu_vec_summary.py gpurun_out/r02_vec_kernels_ncu.csv profiles/r02_vec_kernels_ncu.json"""
import collections, csv, json, sys

n = 4096
N = n * n
nnz = 5 * N - 4 * n
ALG = {  # algorithmic bytes per launch (complex128 = 16 B, kappa f64 = 8 B)
    "hp_stencil_matvec_kernel": (40 * N, "x 16 r + y 16 w + kappa 8 r per grid point"),
    "hp_assemble_csr_kernel": (20 * nnz + 4 * N + 8 * N, "20 B per nonzero (value + column) + 4 B row pointer + 8 B kappa per row"),
    "hp_reduce_kernel<1>": (16 * N, "norm: one vector read"),
    "hp_reduce_kernel<0>": (32 * N, "dot: two vectors read"),
    "hp_axpy_reduce_kernel<0>": (64 * N, "fused w -= h v_j and v_{j+1}.w: v_j, v_{j+1}, w read, w written"),
    "hp_axpy_reduce_kernel<1>": (48 * N, "fused w -= h v_j and |w|: v_j, w read, w written"),
    "hp_combine_kernel": (16 * N * 22, "y = sum of 20 basis vectors: 20 reads + y read + y write"),
    "hp_scale_copy_kernel": (32 * N, "one read, one write"),
}
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    agg.setdefault((r["ID"], name, r["Grid Size"], r["Block Size"]), {})[r["Metric Name"]] = float(r["Metric Value"])
per = collections.OrderedDict()
for (_, name, grid, block), m in agg.items():
    per.setdefault(name, {"grid": grid, "block": block, "launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    p = per[name]
    p["launches"] += 1
    p["ns"] += m["gpu__time_duration.sum"]
    p["rd"] += m["dram__bytes_read.sum"]
    p["wr"] += m["dram__bytes_write.sum"]
out = {"what": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
               "python tools/ncu_vec.py (4096^2, inputs larger than L2); averages per launch",
       "hbm_peak_gbs": json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs"), "kernels": []}
for name, p in per.items():
    L = p["launches"]
    alg, what = ALG.get(name, (None, ""))
    t = p["ns"] / L * 1e-9
    e = {"kernel": name, "grid": p["grid"], "block": p["block"], "launches": L, "duration_us": round(t * 1e6, 2),
         "dram_read_bytes": int(p["rd"] / L), "dram_write_bytes": int(p["wr"] / L), "algorithmic_bytes": alg,
         "algorithmic": what}
    if alg:
        e["traffic_over_algorithmic"] = round((p["rd"] + p["wr"]) / L / alg, 3)
        e["algorithmic_GBs"] = round(alg / t / 1e9, 1)
        if out["hbm_peak_gbs"]:
            e["frac_of_hbm_peak"] = round(alg / t / 1e9 / out["hbm_peak_gbs"], 3)
    out["kernels"].append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for e in out["kernels"]:
    print(e["kernel"], e["duration_us"], e.get("traffic_over_algorithmic"), e.get("frac_of_hbm_peak"))
