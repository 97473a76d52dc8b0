"""SASS instruction counts per kernel family of the shipped library -> profiles/r02_sass_counts.txt
usage: python scratch/sass_counts.py [lib] > profiles/r02_sass_counts.txt   (cuobjdump -sass, sm_100a)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "zisafvm_b200/lib/libzfvm_b200.so"
MNEMONICS = ["LDGSTS", "LDGDEPBAR", "DEPBAR", "UBLKCP", "UBLKPF", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "LDS",
             "STS", "LDG", "STG", "LDL", "STL", "SHFL", "ATOM", "RED", "UTMALDG", "UTCHMMA", "LDTM", "BAR.SYNC"]
FAMILIES = ["flux_kernel", "flux_face_kernel", "update_kernel", "pack", "frozen", "axpy", "cfl", "lsq_weights_kernel",
            "stencil_search_kernel", "source_kernel", "recon_coop_kernel", "recon_tile_kernel", "eq_face_kernel",
            "eq_member_tile_smem_kernel", "eq_member_tile_kernel", "eq_member_kernel", "eq_solve_kernel", "eq_decide_kernel",
            "tracer_generic_kernel", "recon_generic_kernel", "tracer_recon_kernel", "tracer_update_kernel"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts = collections.defaultdict(collections.Counter)
n_inst = collections.Counter()
fam = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        fam = next((f for f in sorted(FAMILIES, key=len, reverse=True) if f in name), "other")
        n_inst[fam] += 1
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and fam:
        op = m.group(1)
        counts[fam]["instr"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[fam][mn] += 1
print(f"SASS of {lib} (cuobjdump -sass), instruction counts per kernel family (scratch/sass_counts.py), end of round 2")
print("mnemonics: LDGSTS = cp.async (global -> shared), LDGDEPBAR/DEPBAR = cp.async commit / wait_group, UBLKCP = cp.async.bulk (TMA bulk),")
print("UBLKPF = cp.async.bulk.prefetch.L2 (K1's prefetch behind the ring), SYNCS = mbarrier, DFMA/DMUL/DADD = FP64 pipe,")
print("MUFU.RSQ64H/RCP64H = fast reciprocal (square root) seeds, ATOM/RED = atomics, UTMALDG = tensor TMA,")
print("UTCHMMA/LDTM = tcgen05 (none: FP64 path, see DESIGN.md section 3)")
print()
print("architectures in the fatbin:", ", ".join(archs))
print()
hdr = f"{'kernel family':28s}{'inst.':>6s}{'instr':>9s}" + "".join(f"{m[:9]:>10s}" for m in MNEMONICS)
print(hdr)
tot = collections.Counter()
for f in list(dict.fromkeys(FAMILIES + ["other"])):
    if not n_inst[f]:
        continue
    print(f"{f:28s}{n_inst[f]:6d}{counts[f]['instr']:9d}" + "".join(f"{counts[f][m]:10d}" for m in MNEMONICS))
    tot.update(counts[f])
print(f"{'total':28s}{'':6s}{tot['instr']:9d}" + "".join(f"{tot[m]:10d}" for m in MNEMONICS))
