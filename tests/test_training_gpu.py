"""SURVEY.md 8f ranks 2-3 on a B200: duration_loss from the integer durations, clip_grad_value_ and
train_step without host synchronisations -- against the reference's own expressions
(glow_tts_train/utils.py:26-28, :118-132; train.py:91-162)."""
from __future__ import annotations

import importlib
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ref_duration_loss(logw, logw_, lengths):             # utils.py:26-28
    return torch.sum((logw - logw_) ** 2) / torch.sum(lengths)


def ref_clip_grad_value_(parameters, clip_value):        # utils.py:118-132
    total = 0.0
    for p in parameters:
        if p.grad is None:
            continue
        total += p.grad.data.norm(2.0).item() ** 2
        p.grad.data.clamp_(min=-clip_value, max=clip_value)
    return total ** 0.5


@pytest.mark.parametrize("shape", [(32, 200), (3, 17), (1, 1), (64, 400)])
def test_duration_loss_matches_the_reference_expression(pkg, shape):
    B, T_x = shape
    g = torch.Generator().manual_seed(B * 1000 + T_x)
    x_len = torch.randint(1, T_x + 1, (B,), generator=g)
    x_len[0] = T_x
    mask = (torch.arange(T_x)[None] < x_len[:, None]).float()
    dur = (torch.randint(0, 12, (B, T_x), generator=g) * mask).to(torch.int32).to(DEV)
    logw = (torch.randn(B, 1, T_x, generator=g) * mask[:, None]).to(DEV).requires_grad_(True)
    # the reference's target, models.py:393, in fp64
    logw_ = (torch.log(1e-8 + dur.double()) * mask.to(DEV).double())[:, None]
    want = ref_duration_loss(logw.double(), logw_, x_len.to(DEV).double())
    (gwant,) = torch.autograd.grad(want, logw)
    got = pkg.duration_loss(logw, dur, x_len)
    (ggot,) = torch.autograd.grad(got, logw)
    assert abs(float(got) - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    assert torch.allclose(ggot.double(), gwant.double(), rtol=1e-5, atol=1e-7)


def test_clip_grad_value_matches_the_reference_and_never_synchronises(pkg):
    g = torch.Generator().manual_seed(3)
    shapes = [(192, 768, 3), (768,), (1, 1), (100000,), (513, 7), (4, 4, 1, 1), (70001,)]
    a = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes] + [torch.nn.Parameter(torch.zeros(3, device=DEV))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a[:-1], b[:-1]):                    # the last parameter has no gradient
        grad = (10 * torch.randn(pa.shape, generator=g)).to(DEV)
        grad.view(-1)[0] = float("nan") if pa.numel() == 513 * 7 else grad.view(-1)[0]
        pa.grad, pb.grad = grad.clone(), grad.clone()
    # (NaN in one gradient: torch.clamp_ keeps it, and so must we; the norm is then NaN in both)
    want = ref_clip_grad_value_(b, 5.0)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")               # any host synchronisation raises
    try:
        got = pkg.clip_grad_value_(a, 5.0)
        got2 = pkg.clip_grad_value_(a, 5.0)               # second call: cached table, already clamped values
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert got.is_cuda and got.dim() == 0
    assert math.isnan(want) and math.isnan(float(got))
    for pa, pb in zip(a[:-1], b[:-1]):
        assert torch.equal(torch.nan_to_num(pa.grad, nan=123.0), torch.nan_to_num(pb.grad, nan=123.0))
    assert math.isnan(float(got2))
    # and without the NaN: the norm itself
    for p in a[:-1]:
        p.grad = torch.nan_to_num((3 * torch.randn(p.shape, generator=g)).to(DEV))
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a[:-1], b[:-1]):
        pb.grad = pa.grad.clone()
    want = ref_clip_grad_value_(b, 1.5)
    got = pkg.clip_grad_value_(a, 1.5)
    assert abs(float(got) - want) <= 1e-6 * want
    for pa, pb in zip(a[:-1], b[:-1]):
        assert torch.equal(pa.grad, pb.grad)


def test_train_step_without_syncs_equals_the_reference_train_step(pkg, oracle):
    """Three optimizer steps from identical weights: the reference's train_step (its own
    clip_grad_value_ and .item() calls) and this package's -- same parameters afterwards, bit for bit."""
    rm = importlib.import_module(oracle.__name__ + ".ref_model")
    ref = rm.import_reference()
    if ref is None:
        pytest.skip("the reference package was not staged (oracle/_ref/refpkg.zip)")
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    train = importlib.import_module(ref.__name__ + ".train")
    batches = [tuple(t if t is None else t.cpu() for t in rm.synthetic_batch(4, 40, 200, seed=s, device="cpu")) for s in (1, 2, 3)]
    prev = rm.swap_monotonic_align(ref, pkg.monotonic_align)
    try:
        results = []
        for step_fn in (train.train_step, pkg.train_step):
            config, model, optimizer = rm.make_model(ref, device=DEV, seed=99)
            torch.manual_seed(5)
            assert step_fn(0, 0, model, optimizer, config, batches, fp16_run=False) == 3
            results.append({n: p.detach().clone() for n, p in model.named_parameters()})
    finally:
        rm.swap_monotonic_align(ref, prev)
    for n in results[0]:
        assert torch.equal(results[0][n], results[1][n]), n
