"""Path consumers at C2 (B=32, D=80, 200 x 1000): the reference's dense-path program (models.py:383-392 +
utils.py:14-23, forward + backward through torch) against the token-level kernels:
  python profiles/time_consumers.py"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
dev = torch.device("cuda:0")
B, D, T_x, T_y = 32, 80, 200, 1000
g = torch.Generator().manual_seed(3)
value = (10 * torch.randn(B, T_x, T_y, generator=g) - 100).to(dev)
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
path, dur, tok = pkg.maximum_path_from_lengths(value, tx, ty, want_durations=True, want_frame_token=True)
x_m = torch.randn(B, D, T_x, generator=g).to(dev).requires_grad_(True)
x_logs = (0.3 * torch.randn(B, D, T_x, generator=g)).to(dev).requires_grad_(True)
z = torch.randn(B, D, T_y, generator=g).to(dev).requires_grad_(True)
logdet = torch.randn(B, generator=g).to(dev).requires_grad_(True)
z_mask = torch.ones(B, 1, T_y, device=dev)
attn = path.unsqueeze(1)
torch.backends.cuda.matmul.allow_tf32 = False


def reference():
    z_m = torch.matmul(attn.squeeze(1).transpose(1, 2), x_m.transpose(1, 2)).transpose(1, 2)
    z_logs = torch.matmul(attn.squeeze(1).transpose(1, 2), x_logs.transpose(1, 2)).transpose(1, 2)
    loss = torch.sum(z_logs) + 0.5 * torch.sum(torch.exp(-2 * z_logs) * ((z - z_m) ** 2))
    loss = loss - torch.sum(logdet)
    loss = loss / torch.sum(torch.ones_like(z) * z_mask) + 0.5 * math.log(2 * math.pi)
    return torch.autograd.grad(loss, [z, x_m, x_logs, logdet])


def ours():
    loss = pkg.aligned_mle_loss(z, x_m, x_logs, logdet, tok, dur, ty)
    return torch.autograd.grad(loss, [z, x_m, x_logs, logdet])


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


a, b = reference(), ours()
for name, r, o in zip(["dz", "dx_m", "dx_logs", "dlogdet"], a, b):
    print(f"{name:8s} max |diff| {(r - o).abs().max().item():.3e}  (max |ref| {r.abs().max().item():.3e})")
print(f"reference: expand by matmul + mle_loss, forward + backward : {timeit(reference):8.1f} us")
print(f"token-level loss kernels, forward + backward               : {timeit(ours):8.1f} us")
