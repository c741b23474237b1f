"""BASELINE.json configs[4] / SURVEY.md 8d (C5), section 4 test (iv): the drop-in inside the LIVE
reference model on a B200.  The reference's own FlowGenerator (models.py) and train_step
(train.py:91-162), packed unmodified into oracle/_ref/refpkg.zip by oracle/build_ref.py, run once with
their own `monotonic_align` (device sync, D2H, the compiled OpenMP Cython kernel, H2D) and once with
`glow_tts_train.models.monotonic_align` replaced by this repository's module -- the two-line swap of
INTEGRATION.md.  Same seed, same inputs: the alignment, the durations, the loss and the gradients
must be IDENTICAL (fp32), also under autocast and for the multi-speaker model (gin_channels=256)."""
from __future__ import annotations

import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref(oracle):
    rm = importlib.import_module(oracle.__name__ + ".ref_model")
    pkg = rm.import_reference()
    if pkg is None:
        pytest.skip("the reference package was not staged (oracle/_ref/refpkg.zip: build in the container that has /root/reference)")
    return rm, pkg


def forward_backward(rm, pkg, model, batch, module, autocast):
    prev = rm.swap_monotonic_align(pkg, module)
    try:
        model.train()
        model.zero_grad(set_to_none=True)
        torch.manual_seed(4321)                       # dropout masks
        utils = importlib.import_module(pkg.__name__ + ".utils")
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            (z, z_m, z_logs, logdet, z_mask), _, (attn, logw, logw_) = model(batch[0], batch[1], batch[2], batch[3], g=batch[4])
            loss = utils.mle_loss(z, z_m, z_logs, logdet, z_mask) + utils.duration_loss(logw, logw_, batch[1])   # train.py:124-129
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        return attn.detach().clone(), logw_.detach().clone(), loss.detach().clone(), grads
    finally:
        rm.swap_monotonic_align(pkg, prev)


@pytest.mark.parametrize("autocast", [False, True])
@pytest.mark.parametrize("variant", ["mean_only", "general", "multi_speaker"])
def test_flow_generator_with_the_module_swapped(ref, pkg, variant, autocast):
    rm, ref_pkg = ref
    kw = {"mean_only": dict(mean_only=True), "general": dict(mean_only=False),
          "multi_speaker": dict(mean_only=True, n_speakers=4, gin_channels=256)}[variant]
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    _, model, _ = rm.make_model(ref_pkg, device=DEV, **kw)
    batch = rm.synthetic_batch(6, 60, 300, n_speakers=kw.get("n_speakers", 1), seed=7, device=DEV)
    theirs = importlib.import_module(ref_pkg.__name__ + ".monotonic_align")
    a = forward_backward(rm, ref_pkg, model, batch, theirs, autocast)
    b = forward_backward(rm, ref_pkg, model, batch, pkg.monotonic_align, autocast)
    assert a[0].dtype == b[0].dtype and a[0].device == b[0].device and a[0].shape == b[0].shape
    assert torch.equal(a[0], b[0]), "attn differs"                      # models.py:378-382
    assert torch.equal(a[1], b[1]), "logw_ differs"                     # models.py:393
    # one 0/1 alignment in, the same program after it: the loss and every gradient agree to the last
    # bit unless cuDNN/atomics reorder a reduction -- which they do not with the flags above
    assert torch.equal(a[2], b[2]), (float(a[2]), float(b[2]))
    assert a[3].keys() == b[3].keys()
    for n in a[3]:
        assert torch.equal(a[3][n], b[3][n]), n
    assert int(a[0].sum()) == int(batch[3].sum() // 1)                   # one token per valid frame


def test_train_step_with_the_module_swapped(ref, pkg):
    """Three optimizer steps of the reference's train_step (train.py:91-162) from identical
    initial weights: identical parameters afterwards."""
    rm, ref_pkg = ref
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    train = importlib.import_module(ref_pkg.__name__ + ".train")
    theirs = importlib.import_module(ref_pkg.__name__ + ".monotonic_align")
    batches = [tuple(t if t is None else t.cpu() for t in rm.synthetic_batch(4, 40, 200, seed=s, device="cpu")) for s in (1, 2, 3)]
    results = []
    for module in (theirs, pkg.monotonic_align):
        config, model, optimizer = rm.make_model(ref_pkg, device=DEV, seed=99)
        prev = rm.swap_monotonic_align(ref_pkg, module)
        try:
            torch.manual_seed(5)
            step = train.train_step(0, 0, model, optimizer, config, batches, fp16_run=False)
        finally:
            rm.swap_monotonic_align(ref_pkg, prev)
        assert step == 3
        results.append({n: p.detach().clone() for n, p in model.named_parameters()})
    for n in results[0]:
        assert torch.equal(results[0][n], results[1][n]), n
