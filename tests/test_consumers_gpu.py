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
