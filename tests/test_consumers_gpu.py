"""SURVEY.md 8(f) rank 1 -- the path's consumers (models.py:383-393) from frame_token / durations:
identical to the reference's dense-path matmuls (forward exactly, backward to fp32 summation noise)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import ragged_lengths

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_consumers(attn, x_m, x_logs, x_mask):
    """models.py:383-393 verbatim (attn: [b,1,t,t'])."""
    z_m = torch.matmul(attn.squeeze(1).transpose(1, 2), x_m.transpose(1, 2)).transpose(1, 2)       # :383-387
    z_logs = torch.matmul(attn.squeeze(1).transpose(1, 2), x_logs.transpose(1, 2)).transpose(1, 2)  # :388-392
    logw_ = torch.log(1e-8 + torch.sum(attn, -1)) * x_mask                                          # :393
    return z_m, z_logs, logw_


@pytest.mark.parametrize("shape", [(3, 80, 37, 150), (4, 80, 200, 1000), (2, 16, 5, 9)])
def test_consumers_match_reference_matmuls(pkg, shape):
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(B * 1000 + T_x)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)).to(DEV)
    tx_d, ty_d = torch.from_numpy(t_x).to(DEV), torch.from_numpy(t_y).to(DEV)
    path, dur, tok = pkg.maximum_path_from_lengths(value, tx_d, ty_d, want_durations=True, want_frame_token=True)
    x_mask = (torch.arange(T_x, device=DEV)[None] < tx_d[:, None]).float().unsqueeze(1)
    x_m = (torch.randn(B, D, T_x, device=DEV) * x_mask).requires_grad_(True)
    x_logs = (torch.randn(B, D, T_x, device=DEV) * x_mask).requires_grad_(True)

    torch.backends.cuda.matmul.allow_tf32 = False
    z_m_ref, z_logs_ref, logw_ref = reference_consumers(path.unsqueeze(1), x_m, x_logs, x_mask)
    z_m = pkg.expand_prior(x_m, tok, dur)
    z_logs = pkg.expand_prior(x_logs, tok, dur)
    logw = pkg.log_durations(dur, tx_d)
    assert torch.equal(z_m, z_m_ref) and torch.equal(z_logs, z_logs_ref)
    assert torch.allclose(logw, logw_ref, rtol=0, atol=1e-6)
    assert logw.shape == logw_ref.shape

    g = torch.randn_like(z_m_ref)
    (gx_ref,) = torch.autograd.grad(z_m_ref, x_m, g, retain_graph=True)
    (gx,) = torch.autograd.grad(z_m, x_m, g)
    assert torch.allclose(gx, gx_ref, rtol=1e-5, atol=1e-5)
    # tokens beyond t_x and frames beyond t_y carry nothing
    for b in range(B):
        assert gx[b, :, t_x[b]:].abs().sum() == 0
        assert z_m[b, :, t_y[b]:].abs().sum() == 0


def reference_mle_loss(z, m, logs, logdet, mask):
    """utils.py:14-23 verbatim (in whatever precision the arguments have)."""
    import math
    loss = torch.sum(logs) + 0.5 * torch.sum(torch.exp(-2 * logs) * ((z - m) ** 2))
    loss = loss - torch.sum(logdet)
    loss = loss / torch.sum(torch.ones_like(z) * mask)
    loss = loss + 0.5 * math.log(2 * math.pi)
    return loss


@pytest.mark.parametrize("mean_only", [False, True])
@pytest.mark.parametrize("shape", [(3, 80, 37, 150), (4, 80, 200, 1000), (2, 16, 5, 9), (1, 7, 1, 3)])
def test_aligned_mle_loss_matches_reference(pkg, shape, mean_only):
    """the loss and its gradients from the token-level prior equal the
    reference's mle_loss on the expanded prior (evaluated in fp64 through the reference's own
    matmul expansion and autograd)."""
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(B * 977 + T_x + int(mean_only))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)).to(DEV)
    tx_d, ty_d = torch.from_numpy(t_x).to(DEV), torch.from_numpy(t_y).to(DEV)
    path, dur, tok = pkg.maximum_path_from_lengths(value, tx_d, ty_d, want_durations=True, want_frame_token=True)
    x_mask = (torch.arange(T_x, device=DEV)[None] < tx_d[:, None]).float().unsqueeze(1)
    z_mask = (torch.arange(T_y, device=DEV)[None] < ty_d[:, None]).float().unsqueeze(1)
    x_m = (torch.randn(B, D, T_x, device=DEV) * x_mask).requires_grad_(True)
    x_logs = None if mean_only else ((0.3 * torch.randn(B, D, T_x, device=DEV) - 0.2) * x_mask).requires_grad_(True)
    z = (torch.randn(B, D, T_y, device=DEV) * z_mask).requires_grad_(True)
    logdet = torch.randn(B, device=DEV).requires_grad_(True)

    loss = pkg.aligned_mle_loss(z, x_m, x_logs, logdet, tok, dur, ty_d)
    inputs = [z, x_m, logdet] + ([] if mean_only else [x_logs])
    grads = torch.autograd.grad(loss, inputs)

    # the reference in fp64: dense-path matmuls (models.py:383-392) + mle_loss (utils.py:14-23)
    z64, xm64, ld64 = (t.detach().double().requires_grad_(True) for t in (z, x_m, logdet))
    xl64 = (torch.zeros_like(xm64) if mean_only else x_logs.detach().double()).requires_grad_(True)
    attn = path.double().unsqueeze(1)
    zm_ref, zl_ref, _ = reference_consumers(attn, xm64, xl64, x_mask.double())
    loss_ref = reference_mle_loss(z64, zm_ref, zl_ref, ld64, z_mask.double())
    grads_ref = torch.autograd.grad(loss_ref, [z64, xm64, ld64] + ([] if mean_only else [xl64]))

    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    for name, g, r in zip(["z", "x_m", "logdet", "x_logs"], grads, grads_ref):
        r = r.float()
        assert g.shape == r.shape, name
        assert torch.allclose(g, r, rtol=1e-4, atol=1e-6 * float(r.abs().max()) + 1e-12), (name, (g - r).abs().max().item())
    # deterministic: a second evaluation gives the same bits
    loss2 = pkg.aligned_mle_loss(z, x_m, x_logs, logdet, tok, dur, ty_d)
    assert torch.equal(loss, loss2)


def reference_generate_path(duration, mask):
    """utils.py:99-115 verbatim (sequence_mask :52-56, convert_pad_shape :41-44 inlined)."""
    import torch.nn.functional as F
    b, t_x, t_y = mask.shape
    cum_duration = torch.cumsum(duration, 1)
    cum_duration_flat = cum_duration.view(b * t_x)
    x = torch.arange(t_y, dtype=cum_duration_flat.dtype, device=duration.device)
    path = (x.unsqueeze(0) < cum_duration_flat.unsqueeze(1)).to(mask.dtype)
    path = path.view(b, t_x, t_y)
    path = path - F.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]
    path = path * mask
    return path


@pytest.mark.parametrize("shape", [(3, 37, 150), (4, 200, 1000), (1, 1, 1), (2, 64, 50)])
def test_generate_path_matches_reference(pkg, shape):
    """the inference-side durations -> path expansion (models.py:340)."""
    B, T_x, T_y = shape
    rng = np.random.default_rng(B * 31 + T_x)
    t_x, _ = ragged_lengths(rng, B, T_x, max(T_x, T_y))
    x_mask = (torch.arange(T_x, device=DEV)[None] < torch.from_numpy(t_x).to(DEV)[:, None]).float()
    # ceil-ed durations as models.py:330-333 makes them; the frame budget comes from their sum (:334)
    w_ceil = torch.ceil(torch.rand(B, T_x, device=DEV) * 2 * T_y / T_x) * x_mask
    y_len = torch.clamp_min(w_ceil.sum(1), 1).long()
    y_mask = (torch.arange(T_y, device=DEV)[None] < y_len[:, None]).float()
    attn_mask = (x_mask.unsqueeze(-1) * y_mask.unsqueeze(1)).unsqueeze(1)          # [b,1,t_x,t_y], models.py:334-337
    want = reference_generate_path(w_ceil, attn_mask.squeeze(1))
    got = pkg.generate_path(w_ceil, attn_mask.squeeze(1))
    assert got.dtype == want.dtype and got.shape == want.shape
    assert torch.equal(got, want)
    # a strided mask view and integer durations give the same path
    wide = torch.zeros(B, T_x, T_y + 5, device=DEV)
    wide[:, :, :T_y] = attn_mask.squeeze(1)
    assert torch.equal(pkg.generate_path(w_ceil.long(), wide[:, :, :T_y]), want)
