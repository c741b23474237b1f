"""Re-measure what systolic::choose_shape (mas_dp_cta.cuh) hard-codes: the cost of one 32-frame step
of kernel (1) per tokens-per-lane R, on THIS device.

    python profiles/measure_step_costs.py            # prints a table to paste into kStep[]

For each R the kernel is forced (MAS_B200_FORCE_R, read once per process: one subprocess per R) on a
latency-bound batch (B=32, 192 x 2048: one CTA per SM or fewer, the step is what is timed), and the
time per launch is turned into cycles per (blocks + W - 1) steps at the SM clock nvidia-smi reports."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import __graft_entry__ as entry
pkg = entry.load_package()
B, T_x, T_y = 32, 192, 2048
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
vals = [(10 * torch.randn(B, T_x, T_y, generator=g) - 100).to(dev) for _ in range(6)]
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev); ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for i in range(3): pkg.maximum_path_from_lengths(vals[i], tx, ty)
graph = torch.cuda.CUDAGraph(); keep = []
with torch.cuda.graph(graph):
    for i in range(12): keep.append(pkg.maximum_path_from_lengths(vals[i %% 6], tx, ty))
graph.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
print(e0.elapsed_time(e1) * 1e3 / 12)
''' % str(ROOT)


def main():
    try:
        mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True).stdout.split()[0])
    except (OSError, IndexError, ValueError):
        mhz = 1965.0
    print(f"SM clock {mhz:.0f} MHz; B=32, 192 x 2048 (64 blocks); cycles per step = us * MHz / (64 + W - 1), backtrack and output included")
    for R in (1, 2, 3, 4, 5, 6, 8):
        env = dict(os.environ, MAS_B200_FORCE_R=str(R))
        out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env)
        if out.returncode != 0:
            print(f"R={R}: failed ({out.stderr.strip().splitlines()[-1] if out.stderr.strip() else 'no output'})")
            continue
        us = float(out.stdout.split()[-1])
        W = -(-6 // R)
        print(f"R={R} W={W}: {us:8.1f} us per launch -> {us * mhz / (64 + W - 1):7.0f} cycles per step")


if __name__ == "__main__":
    main()
