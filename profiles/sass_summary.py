"""Mnemonic counts per kernel from `cuobjdump -sass` of the objects linked into libmas_b200.so
(run here, no GPU needed):  python profiles/sass_summary.py   ->  profiles/r2_sass_{fused,systolic,logp}.txt"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "glow-tts-train_b200" / "build"
KEYS = ["UTMALDG", "UBLKCP", "UTMACCTL", "SYNCS", "UCGABAR", "FFMA2", "FMUL2", "FFMA", "FMNMX", "SHFL.UP", "LDS.128", "STS.128",
        "STAS", "LDGSTS", "BAR.SYNC", "MEMBAR", "CCTL", "STL", "LDL", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA"]
JOBS = [("fused", "mas_fused.o", "mas_fused_kernel", "Kernel (2) mas_fused_kernel<R, dbg>"),
        ("systolic", "mas_path_systolic.o", "mas_path_systolic_kernel", "Kernel (1) mas_path_systolic_kernel<R, threads, dbg, cluster>"),
        ("logp", "mas_logp.o", "mas_logp_kernel", "mas_logp_kernel")]
for tag, obj, name, title in JOBS:
    sass = subprocess.run(["cuobjdump", "-sass", str(BUILD / obj)], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1) if name in m.group(1) else None
            if cur:
                per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            per[cur]["instructions"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or op.startswith(k + "_"):
                    per[cur][k] += 1
                    break
    with open(ROOT / "profiles" / f"r2_sass_{tag}.txt", "w") as f:
        f.write(f"{title}: SASS summary (round 2)\n\n")
        f.write("mnemonic counts per kernel (cuobjdump -sass of the object linked into libmas_b200.so, sm_100a only;\n"
                "`python profiles/sass_summary.py`; no UTC*MMA / LDTM / HMMA anywhere: the tensor cores are unused by decision)\n")
        for fn, c in per.items():
            f.write(fn + "\n    instructions " + str(c["instructions"]) + ": " +
                    ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]) + "\n")
    print(tag, len(per), "kernels")
