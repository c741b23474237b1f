"""Evidence for the north star's "no tensor cores: K = 80 is too thin" (VERDICT r1, next #7).

The C2 contraction -- per utterance [200 x 160] . [160 x 1000] (the inv_var and the mean term side by
side), fp32 accumulate -- on the tensor cores through cuBLAS, in the three precisions a tensor-core
kernel could offer, against this repository's FFMA2 loop:

  fp32      torch.bmm, float32 matmul precision "highest" (cuBLAS' own fp32 path)
  tf32      one TF32 pass (allow_tf32): what a plain tcgen05 kind::tf32 kernel computes
  3xtf32    split TF32: a = a_hi + a_lo (a_hi = a rounded to TF32), c = a_hi b_hi + a_hi b_lo + a_lo b_hi,
            three TF32 passes, fp32 accumulate -- the error-compensated variant

For each: time per batch of the contraction ALONE (CUDA events, median of 20; the element-wise
operand preparation and the final adds are not timed -- they are the same for every variant and our
kernel does them inside its loop), max relative error of logp against the fp64 formula (tolerance of
the north star: 1e-5), and how many frames of the final alignment sit on a different token than with
the fp64 scores (kernel (1) run on each matrix).  Library GEMMs, not a hand-written tcgen05 kernel:
the point is what the ARITHMETIC and the hardware's GEMM rate give at K = 80 + 80, not a tuned kernel.

    python profiles/probes/probe_tf32.py          # on a B200
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
oracle = entry.load_oracle()
dev = torch.device("cuda:0")
NEG_HALF_LOG_2PI = -0.5 * np.log(2 * np.pi)


def tf32_round(x):
    """Round-to-nearest-even onto TF32's 10 mantissa bits (what cvt.rna.tf32.f32 gives up to tie handling)."""
    i = x.view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def operands(x_m, x_logs, z):
    r = torch.exp(-2 * x_logs)
    a = torch.cat([-0.5 * r, x_m * r], 1).transpose(1, 2).contiguous()      # [B, T_x, 2D]
    b = torch.cat([z * z, z], 1).contiguous()                               # [B, 2D, T_y]
    l1 = (NEG_HALF_LOG_2PI - x_logs).sum(1)                                 # [B, T_x]
    l4 = (-0.5 * x_m * x_m * r).sum(1)
    return a, b, l1, l4


def contraction(kind, a, b):
    if kind == "fp32":
        torch.backends.cuda.matmul.allow_tf32 = False
        return torch.bmm(a, b)
    torch.backends.cuda.matmul.allow_tf32 = True
    if kind == "tf32":
        return torch.bmm(a, b)
    a_hi, b_hi = tf32_round(a), tf32_round(b)
    a_lo, b_lo = a - a_hi, b - b_hi
    return torch.bmm(a_hi, b_hi) + (torch.bmm(a_hi, b_lo) + torch.bmm(a_lo, b_hi))


def time_us(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    B, D, T_x, T_y = 32, 80, 200, 1000
    g = torch.Generator().manual_seed(1234)
    x_m = torch.randn(B, D, T_x, generator=g)
    x_logs = 0.3 * torch.randn(B, D, T_x, generator=g) - 0.5
    idx = (torch.arange(T_y) * T_x) // T_y
    z = x_m[:, :, idx] + torch.exp(x_logs[:, :, idx]) * torch.randn(B, D, T_y, generator=g)
    ref64 = oracle.logp_f64(x_m.numpy(), x_logs.numpy(), z.numpy())
    t_x, t_y = np.full(B, T_x, np.int32), np.full(B, T_y, np.int32)
    want = oracle.maximum_path(ref64.astype(np.float32), t_x, t_y, threads=8)
    xm, xl, zz = x_m.to(dev), x_logs.to(dev), z.to(dev)
    tx, ty = torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev)
    a, b, l1, l4 = operands(xm, xl, zz)
    print(f"C2 contraction, B={B}: [{T_x} x {2 * D}] . [{2 * D} x {T_y}] per utterance, 2.05 GFLOP")
    print(f"{'variant':10s} {'us / batch':>11s} {'max rel err vs fp64':>20s} {'frames moved':>13s}")

    def report(name, us, logp):
        lp = logp.cpu().numpy()
        rel = float(np.max(np.abs(lp - ref64) / np.abs(ref64)))
        path = pkg.maximum_path_from_lengths(logp.contiguous(), tx, ty).cpu().numpy().astype(np.int32)
        moved = int((path != want).sum() // 2)
        print(f"{name:10s} {us:11.1f} {rel:20.2e} {moved:13d}")

    ours = lambda: pkg.log_likelihood_matrix(xm, xl, zz)  # noqa: E731
    report("ffma2", time_us(ours), ours())
    for kind in ("fp32", "tf32", "3xtf32"):
        fn = lambda: contraction(kind, a, b)  # noqa: E731
        c = fn()
        report(kind, time_us(fn), (l1[:, :, None] + c) + l4[:, :, None])
    print("(ffma2 = this repository's kernel, operand preparation and final adds included; the cuBLAS rows time the GEMMs only)")


if __name__ == "__main__":
    main()
