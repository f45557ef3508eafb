"""SASS evidence per kernel from the shipped objects (here, no GPU needed):  python profiles/sass_summary.py > profiles/sass_rNN_summary.txt
Counts the mnemonics that prove the Blackwell path (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = bulk async copy) and the
wide memory instructions (REDG .F32x4 vector reductions, LDG.E.ENL2.256 256-bit gathers)."""
import collections
import re
import subprocess

KERNELS = (("field_mlp", ["mlp_rows_gemm_kernel", "mlp_wgrad_kernel", "mlp_pack_weights_kernel"]),
           ("gbuffer", ["gb_bwd_kernel", "gb_bwd_finalize_xfm_kernel", "gb_fwd_kernel"]),
           ("antialias", ["aa_bwd_pair_kernel", "aa_fwd_pair_kernel"]), ("marching_tets", ["mt_tcount_kernel"]), ("normals", ["normals_splat_kernel"]))
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "SYNCS", "FENCE", "REDG", "LDG", "STG", "LDS", "STS", "ATOMG", "HMMA", "FFMA", "FMUL", "MUFU"]
print("# SASS evidence (cuobjdump -sass of the shipped objects in 3danimals_b200/csrc/_build, sm_100a only)\n")
for obj, kernels in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", "3danimals_b200/csrc/_build/%s.o" % obj], capture_output=True, text=True).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", txt))
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0]
        if not any(k in name for k in kernels):
            continue
        ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", f, flags=re.M))
        summ = {}
        for op, c in ops.items():
            for key in KEYS:
                if op.startswith(key):
                    full = op if key in ("REDG", "LDG", "LDTM", "UBLKCP", "UTCHMMA") else key
                    summ[full] = summ.get(full, 0) + c
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:120]
        print("%s.o [%s]  %s\n    %d SASS instructions; %s\n" % (obj, ",".join(sorted(arch)), dem, sum(ops.values()), ", ".join("%s x%d" % kv for kv in sorted(summ.items()))))
