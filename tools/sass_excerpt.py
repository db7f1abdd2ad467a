"""Writes profiles/r2_mlp_tc_sass_excerpt.txt: per production variant of mlp_tc_kernel the counts of the SASS
mnemonics that prove the tcgen05 / tensor-memory / bulk-copy data path (B200_PROFILING.md), plus the first lines
carrying each of them.  Runs where the object exists (build container): cuobjdump -sass evpp_b200/lib/obj/mlp_tc.o."""
import collections
import os
import re
import subprocess


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "evpp_b200", "lib", "obj", "mlp_tc.o")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FENCE.VIEW.ASYNC", "UTCATOMSWS", "F2FP", "FFMA2", "FMNMX", "HMMA", "MUFU"]
VARIANTS = {
    "ILb0ELb1ELi2ELi0ELi2ELi1ELi1E": "bulk fp16, 2-CTA cluster, sigma-only coarse pass (HEADS=1)",
    "ILb0ELb1ELi2ELi0ELi2ELi1ELi3E": "bulk fp16, 2-CTA cluster, beta-only fine pass with the folded feature layer (HEADS=3, the production fine pass)",
    "ILb0ELb1ELi2ELi0ELi2ELi1ELi2E": "bulk fp16, 2-CTA cluster, beta-only fine pass in the reference's layer order (HEADS=2)",
    "ILb0ELb1ELi2ELi0ELi2ELi1ELi0E": "bulk fp16, 2-CTA cluster, full outputs (render())",
    "ILb0ELb1ELi2ELi0ELi4ELi3ELi3E": "exact fp16x3 (3 MMAs per product), 2-CTA cluster, beta-only fine pass with the folded feature layer",
    "ILb0ELb1ELi2ELi0ELi4ELi3ELi1E": "exact fp16x3, 2-CTA cluster, sigma-only coarse pass",
}
sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
out = ["# SASS evidence for mlp_tc_kernel (sm_100a), from: cuobjdump -sass evpp_b200/lib/obj/mlp_tc.o",
       "# UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st (tensor memory), UBLKCP = cp.async.bulk (weight stream,",
       "# .MULTICAST = cluster multicast), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 = packed fp32 FMA (bias + scale),",
       "# F2FP = fp32 -> packed 16-bit conversion (ReLU fused), HMMA = legacy warp-level MMA (must be 0 here)", ""]
funcs = re.split(r"\n\s*Function : ", sass)[1:]
for fn in funcs:
    name, body = fn.split("\n", 1)
    key = next((k for k in VARIANTS if k in name), None)
    if key is None:
        continue
    cnt = collections.Counter()
    first = {}
    n_instr = 0
    for line in body.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
        if not m:
            continue
        n_instr += 1
        ins = m.group(1)
        for w in WANT:
            if re.search(r"(^|\s|@!?U?P\d+\s+)" + re.escape(w), ins):
                cnt[w] += 1
                first.setdefault(w, ins.strip())
        if "UBLKCP" in ins and "MULTICAST" in ins:
            cnt["UBLKCP.MULTICAST"] += 1
    out.append(f"## {VARIANTS[key]}")
    out.append(f"   {name.strip()}")
    out.append(f"   {n_instr} SASS instructions; " + ", ".join(f"{w} {cnt[w]}" for w in WANT + ["UBLKCP.MULTICAST"]))
    for w in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR"):
        if w in first:
            out.append(f"     first {w}: {first[w]}")
    out.append("")
path = os.path.join(ROOT, "profiles", "r2_mlp_tc_sass_excerpt.txt")
open(path, "w").write("\n".join(out))
print("\n".join(out))
