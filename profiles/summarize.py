"""Turn the reports profiles/collect.sh brought back (gpurun_out/) into the tracked text summaries
under profiles/:  python profiles/summarize.py r1"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = Path("profiles")
src = Path("gpurun_out")

# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(src / f"{R}_launches.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
per = collections.defaultdict(list)
for r in rows[1:]:
    per[r[ki]].append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in per.values())
with open(out / f"{R}_launches.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-c5`\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k[:90]:92s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:9.1f} us share={sum(v) / total:6.1%}\n")

# ---- full captures -----------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__cluster", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
traffic = {}
for tag in ("mas", "fused", "logp"):
    rep = src / f"{R}_prof_{tag}.ncu-rep"
    if not rep.exists():
        continue
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(raw.splitlines()))
    h, u, v = rws[0], rws[1], rws[2]
    kname = v[h.index("Kernel Name")]
    with open(out / f"{R}_ncu_{tag}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch of: {kname}\n")
        vals = {}
        for a, b, c in zip(h, u, v):
            vals[a] = (c, b)
            if any(w in a for w in WANT) and ".min" not in a and ".max." not in a:
                f.write(f"{a:88s} {c:>18s} {b}\n")
        stalls = {a: c for a, (c, b) in vals.items() if "warps_issue_stalled" in a and a.endswith("per_issue_active.ratio")}
        f.write("# stall reasons (warps per issue-active cycle)\n")
        for a, c in sorted(stalls.items(), key=lambda kv: -float(kv[1] or 0))[:8]:
            f.write(f"{a:88s} {c:>18s}\n")
    def mb(name):
        c, unit = vals[name]
        x = float(c.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    traffic[tag] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    # where the warps' time goes, by code region: the hottest loop (by executed count) against the rest
    srcpage = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
    srows = list(csv.reader(srcpage.splitlines()))
    if len(srows) > 3 and "Source" in srows[1]:
        sh = srows[1]
        ia, cs, ce = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
        st = [(i, x) for i, x in enumerate(sh) if x.startswith("stall_") and "Not Issued" not in x]
        body = [x for x in srows[2:] if len(x) > ce]
        mx = max(float(x[ce] or 0) for x in body)
        hot = [i for i, x in enumerate(body) if float(x[ce] or 0) > 0.9 * mx]
        lo, hi = min(hot), max(hot)
        tot = sum(float(x[cs] or 0) for x in body)

        def region(a, b):
            d = collections.Counter()
            for x in body[a:b]:
                for i, n in st:
                    d[n[6:]] += float(x[i] or 0)
            return sum(float(x[cs] or 0) for x in body[a:b]), d

        with open(out / f"{R}_ncu_{tag}.txt", "a") as f:
            f.write(f"# warp-state samples by code region ({int(tot)} samples, {len(body)} SASS instructions)\n")
            for name, (a, b) in {"hottest loop": (lo, hi + 1), "before it": (0, lo), "after it": (hi + 1, len(body))}.items():
                n, d = region(a, b)
                top = ", ".join(f"{k} {int(v)}" for k, v in d.most_common(5) if v > 0)
                f.write(f"#   {name:13s} instr {a:5d}..{b - 1:5d}: {n / max(tot, 1):6.1%} of samples ({top})\n")
            mix = collections.Counter((x[ia].split()[1] if x[ia].strip().startswith("@") else x[ia].split()[0]) for x in body[lo:hi + 1])
            f.write("#   hottest loop instruction mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(8)) + "\n")
# (a round that did not re-capture a kernel keeps the previous round's figure for it)
prev = json.load(open(out / "traffic.json")) if (out / "traffic.json").exists() else {}
new = {"c2": traffic.get("fused"), "c1": traffic.get("mas"), "fused_c2": traffic.get("fused"), "logp_c2": traffic.get("logp")}
merged = {k: (v if v is not None else prev.get(k)) for k, v in new.items()}
merged["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, B=32 200x1000"
merged["captured_in"] = {**prev.get("captured_in", {}), **{k: R for k, v in new.items() if v is not None}}
json.dump(merged, open(out / "traffic.json", "w"), indent=1)
print(open(out / f"{R}_launches.txt").read())
