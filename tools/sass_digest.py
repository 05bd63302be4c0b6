"""Instruction-mix digest of the built kernels (cuobjdump -sass of csrc/*.o): the mnemonics that show which hardware
paths a kernel uses (bulk-copy TMA, distributed shared memory, mbarriers, cluster barriers, FP64 FMA, shuffles).
    python tools/sass_digest.py > profiles/r02_sass_digest.md        (build container, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "helmholtz_preconditioner_b200", "csrc")
WATCH = ["UBLKCP", "UBLKPF", "UTMALDG", "UTCHMMA", "UTCQMMA", "STAS", "SYNCS", "UCGABAR", "DFMA", "DMUL", "DADD", "SHFL",
         "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "ATOMG", "RED", "MUFU", "HMMA", "IMMA", "DMMA"]
KERNELS = [("hp_sweep4.o", r"hp_sweep4_kernelILi[012]ELb0ELi12ELi4E"), ("hp_sweep4d.o", r"hp_sweep4d_kernelILi[01]ELb0ELi12ELi4E"),
           ("hp_sweep4m.o", r"hp_sweep4m_kernelILi0ELb0ELi12ELi4ELi8E"), ("hp_peer.o", r"hp_handover_kernel"), ("hp_cgs.o", r"hp_cgs_kernelILi12ELb1ELb1E"),
           ("hp_assembly.o", r"hp_stencil_matvec"), ("hp_assembly.o", r"hp_assemble_csr"), ("hp_blas.o", r"hp_axpy_reduce_kernelILi0E"),
           ("hp_blas.o", r"hp_reduce_kernelILi0E"), ("hp_blas.o", r"hp_combine"), ("hp_front_coupled.o", r"hp_fc_leaf_solve_kernelILi12E"),
           ("hp_setup.o", r"hp_chain_reg_kernelILi12ELb1E"), ("hp_setup.o", r"hp_leaf_warp_kernelILi12E")]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            res[cur][m.group(1).split(".")[0]] += 1
    return res


def main():
    print("# SASS instruction mix (sm_100a), `python tools/sass_digest.py`\n")
    print("Counts of static instructions per kernel.  UBLKCP = cp.async.bulk (TMA 1-D bulk copy), UBLKPF = bulk L2 prefetch, "
          "STAS = st.async to distributed shared memory, SYNCS = mbarrier ops, UCGABAR = cluster barrier, DFMA = FP64 FMA.  "
          "No UTMALDG / UTC*MMA / HMMA is expected: the packets are contiguous 1-D copies and complex128 has no tcgen05 path; the "
          "8-right-hand-side sweep (hp_sweep4d) uses the FP64 tensor-core instruction DMMA (mma.sync.m8n8k4.f64).\n")
    cache = {}
    cols = WATCH
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for obj, pat in KERNELS:
        if not os.path.exists(os.path.join(CSRC, obj)):
            continue
        if obj not in cache:
            cache[obj] = functions(obj)
        for name, cnt in cache[obj].items():
            if re.search(pat, name):
                dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
                dem = dem.replace("void ", "").replace("(int)", "").replace("(bool)", "")
                dem = dem[:dem.index(">(") + 1] if ">(" in dem else dem[:dem.index("(")] if "(" in dem else dem
                pref = lambda c: sum(v for k, v in cnt.items() if k == c or k.startswith(c + "_"))
                print(f"| `{dem}` | {sum(cnt.values())} | " + " | ".join(str(pref(c)) for c in cols) + " |")


if __name__ == "__main__":
    main()
