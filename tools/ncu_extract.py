"""Developer tool: key metrics of every kernel of an .ncu-rep (ncu --page raw --csv) as a small JSON for profiles/.
usage: python tools/ncu_extract.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__cluster_max_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": r[i], "unit": units[i]}
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                st[h[len(STALL):-len("_per_issue_active.ratio")]] = round(v, 2)
    d["stalls_per_issue"] = st
    out.append(d)
json.dump(out if len(out) > 1 else out[0], open(sys.argv[2], "w"), indent=1)
print(f"{len(out)} kernels -> {sys.argv[2]}")
