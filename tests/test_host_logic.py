"""Host-side logic that needs no GPU: the no-fallback guarantee, utterance sharding, and the N>1
data path (world_size-2 gloo run on CPU)."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


def test_no_cpu_fallback(pkg):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    v, m = torch.zeros(1, 3, 5), torch.ones(1, 3, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.monotonic_align.maximum_path(v, m)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pkg.maximum_path_from_lengths(v, torch.ones(1, dtype=torch.int32), torch.ones(1, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pkg.fused_maximum_path(torch.zeros(1, 80, 3), None, torch.zeros(1, 80, 5), torch.tensor([3]), torch.tensor([5]))


def test_product_does_not_import_oracle():
    """The product path may not route through the oracle (or the reference) in any form."""
    for path in (REPO / "glow-tts-train_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = path.read_text()
            assert "oracle" not in text.replace("no oracle", ""), path
            assert "/root/reference" not in text, path


def test_signature_matches_reference(pkg):
    import inspect

    sig = inspect.signature(pkg.monotonic_align.maximum_path)
    names = [n for n, p in sig.parameters.items() if p.kind is p.POSITIONAL_OR_KEYWORD]
    assert names == ["value", "mask"]      # monotonic_align/__init__.py:6


def test_contiguous_shard(pkg):
    from glow_tts_train_b200 import sharding

    for B in (0, 1, 7, 32, 256):
        for ws in (1, 2, 3, 4, 8):
            got = [i for r in range(ws) for i in sharding.contiguous_shard(B, ws, r)]
            assert got == list(range(B))
            sizes = [len(sharding.contiguous_shard(B, ws, r)) for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.contiguous_shard(4, 2, 2)


def test_balanced_shards(pkg):
    from glow_tts_train_b200 import sharding

    rng = np.random.default_rng(0)
    costs = np.sort(rng.integers(1000, 200000, 64))[::-1]
    shards = sharding.balanced_shards(costs.tolist(), 8)
    assert sorted(i for s in shards for i in s) == list(range(64))
    loads = [sum(costs[i] for i in s) for s in shards]
    contiguous = [sum(costs[i] for i in sharding.contiguous_shard(64, 8, r)) for r in range(8)]
    assert max(loads) / min(loads) < 1.1 < max(contiguous) / min(contiguous)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, str(REPO))
    import torch.distributed as dist

    import __graft_entry__ as entry

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    oracle = entry.load_oracle()
    pkg = entry.load_package()
    from glow_tts_train_b200 import sharding

    rng = np.random.default_rng(99)                       # same batch on every rank
    B, T_x, T_y = 7, 12, 40
    value = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
    t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
    t_y = np.array([rng.integers(t, T_y + 1) for t in t_x], np.int32)
    mine = sharding.contiguous_shard(B, world_size, rank)
    # the shard is processed with NO data-path collective; the CPU oracle stands in for the kernel here
    local = oracle.maximum_path(value[mine.start:mine.stop], t_x[mine.start:mine.stop], t_y[mine.start:mine.stop])
    # bench-style reduction: units processed are summed, time is the max over ranks
    cells = torch.tensor([local.size], dtype=torch.float64)
    elapsed = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    gathered = [None] * world_size
    dist.all_gather_object(gathered, (mine.start, local))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])])
        want = oracle.maximum_path(value, t_x, t_y)
        np.save(Path(out_dir) / "ok.npy", np.array([np.array_equal(full, want), cells.item() == value.size,
                                                     abs(elapsed.item() - 0.001 * world_size) < 1e-12]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_gloo(tmp_path, oracle, pkg):
    """world_size 2 on CPU (gloo): the concatenation of per-rank results equals the single-process
    result, cells are summed and time is max-reduced -- the N>1 contract of bench.py."""
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all()


def _deal(lib, P, BT, nchunks):
    owner = np.empty(BT * nchunks, np.int32)
    order = np.empty(BT * nchunks, np.int32)
    rc = lib.mas_b200_debug_deal(P, BT, nchunks, owner.ctypes.data, order.ctypes.data)
    return rc, owner.reshape(BT, nchunks), order.reshape(BT, nchunks)


def test_contraction_units_are_dealt_exactly_once(pkg):
    """Host logic of the contraction kernels (csrc/mas_logp_cta.cuh `Deal`): every (row, chunk) unit goes
    to exactly one persistent CTA, whatever the numbers of CTAs, rows and chunks."""
    lib = pkg._lib.load()
    rng = np.random.default_rng(5)
    combos = [(148, 32, 13), (116, 32, 13), (148, 160, 2), (148, 148, 7), (148, 149, 7), (1, 5, 3), (7, 1, 40),
              (132, 32, 13), (100, 48, 13), (148, 1, 13), (3, 3, 1)]
    combos += [tuple(int(v) for v in (rng.integers(1, 200), rng.integers(1, 400), rng.integers(1, 40))) for _ in range(300)]
    for P, BT, n in combos:
        rc, owner, order = _deal(lib, P, BT, n)
        assert rc == 0, (P, BT, n, rc)
        assert owner.min() >= 0 and owner.max() < P, (P, BT, n)
        # a CTA walks the chunks of a row in ascending order: early frames first
        for r in range(BT):
            for cta in np.unique(owner[r]):
                mine = np.flatnonzero(owner[r] == cta)
                assert np.all(np.diff(order[r, mine]) > 0), (P, BT, n, r, cta)


def test_contraction_deal_at_the_benchmark_shapes(pkg):
    """C2 (32 utterances x 13 chunks): the materialising kernel needs 3 rounds on 148 SMs, the single
    launch 4 on the 116 SMs the sweeps leave, where the spare CTAs have the time to take chunk 9, just
    before the dedicated CTAs' last round; with 3 rounds they take the last chunk."""
    lib = pkg._lib.load()
    for P, rounds in ((148, 3), (116, 4)):
        rc, owner, order = _deal(lib, P, 32, 13)
        assert rc == 0
        per_cta = np.bincount(owner.ravel(), minlength=P)
        dedicated = per_cta[: (P // 32) * 32]
        assert dedicated.max() == rounds and dedicated.min() >= rounds - 1
        spares = per_cta[(P // 32) * 32:]
        assert spares.sum() == 32 and spares.max() <= 2
        spare_chunks = {int(c) for r in range(32) for c in np.flatnonzero(owner[r] >= (P // 32) * 32)}
        assert spare_chunks == ({9} if P == 116 else {12})


def test_contraction_tile_shapes(pkg):
    """csrc/mas_logp_tile.cuh `make_tile_shape`: at most 512 threads of 4 x 8 cells, whole utterance covered."""
    lib = pkg._lib.load()
    out = np.empty(6, np.int32)
    for T_x, T_y in [(200, 1000), (1, 1), (7, 8), (256, 64), (257, 264), (400, 2000), (1024, 8192), (2048, 65536), (50, 122)]:
        assert lib.mas_b200_debug_tile_shape(T_x, T_y, out.ctypes.data) == 0
        row_tiles, tile_rows, RG, CG, F, nchunks = (int(v) for v in out)
        assert RG * CG <= 512 and 8 <= CG <= 16 and F == 8 * CG and tile_rows == 4 * RG and tile_rows <= 256
        assert row_tiles * tile_rows >= T_x and (row_tiles - 1) * tile_rows < T_x
        assert nchunks * F >= T_y and (nchunks - 1) * F < T_y
    lib.mas_b200_debug_tile_shape(200, 1000, out.ctypes.data)
    assert tuple(out) == (1, 200, 50, 10, 80, 13)


def test_sweep_plans(pkg):
    """Host logic of kernel (1) (`choose_shape` / `choose_plan`): tokens covered, shared memory within
    the device's limit, one CTA per utterance unless capacity needs a cluster."""
    lib = pkg._lib.load()
    out = np.empty(8, np.int32)
    max_smem, num_sms = 232448 - 2048, 148                      # B200: opt-in shared memory per CTA, SMs
    want = {
        (32, 200, 1000): (3, 3, 1, 1),                          # R, W, K, bits in shared memory
        (256, 400, 2000): (3, 5, 1, 0),                         # 129 KB of direction bits: in the workspace
        (8, 1024, 8192): (2, 2, 8, 1),                          # long form: 8 CTAs per utterance hold the bits
        (32, 1024, 8192): (3, 3, 4, 0),
        (2, 5, 36): (1, 1, 1, 1),
    }
    rng = np.random.default_rng(9)
    shapes = list(want) + [(int(rng.integers(1, 300)), int(rng.integers(1, 2049)), 0) for _ in range(200)]
    for B, T_x, T_y in shapes:
        if T_y == 0:
            T_y = int(rng.integers(max(T_x, 32), 65537)) // 4 * 4
        rc = lib.mas_b200_debug_path_plan(B, T_x, T_y, max_smem, num_sms, out.ctypes.data)
        assert rc == 0, (B, T_x, T_y, rc)
        R, W, S, K, rows, nblk, bits_smem, total = (int(v) for v in out)
        assert rows == 32 * R * W and rows * K >= T_x and nblk == -(-T_y // 32)
        assert 2 <= S <= 4 and K in (1, 2, 4, 8) and W <= 15 and total <= max_smem
        if (B, T_x, T_y) in want:
            assert (R, W, K, bits_smem) == want[(B, T_x, T_y)], (B, T_x, T_y, out.tolist())


def test_fused_geometry_for_the_benchmark_shapes(pkg):
    """Kernel (2)'s launch geometry is pure host arithmetic (mas_fused.cu: make_geom / choose_geom):
    cluster size, slice, FFMA teams and the shared-memory budget for the shapes bench.py times and for
    random ones, on a B200's limits (148 SMs, 227 KB of opt-in shared memory per CTA)."""
    lib = pkg._lib.load()
    max_smem, num_sms = 232448 - 1024, 148
    out = np.zeros(12, np.int32)
    want = {  # (B, T_x, T_y) -> (CTAs per utterance, tokens per sweep lane, tokens per CTA, teams)
        (32, 200, 1000): (4, 2, 50, 3),      # C2: 128 of 148 SMs, three contraction passes per slice
        (256, 400, 2000): (4, 4, 100, 3),    # C3: 37 clusters at a time, direction bits in shared memory behind a 6-box ring
        (8, 1024, 8192): (8, 4, 128, 3),     # C4: direction bits in the workspace
    }
    rng = np.random.default_rng(11)
    cases = list(want) + [(int(rng.integers(1, 600)), int(rng.integers(1, 1025)), 0) for _ in range(300)]
    for B, T_x, T_y in cases:
        if T_y == 0:
            T_y = int(rng.integers(max(T_x, 4), 8193)) // 4 * 4
        rc = lib.mas_b200_debug_fused_geom(B, 80, T_x, T_y, max_smem, num_sms, out.ctypes.data)
        assert rc == 0, (B, T_x, T_y, rc)
        K, R, max_slice, nteams, team_warps, CG, F, NB, bits_smem, total, ffma_warps, ring_rows = (int(v) for v in out)
        assert K in (1, 2, 4, 8) and 1 <= R <= 8 and max_slice % R == 0
        assert max_slice * K >= T_x and max_slice <= 32 * R and ring_rows >= max_slice and ring_rows % 8 == 0
        assert nteams * team_warps == ffma_warps and ffma_warps in (12, 15)
        assert F == 8 * CG and -(-max_slice // 4) * CG <= 32 * team_warps      # a team's register tiles fit its threads
        assert NB * 32 >= F + 64 and total <= max_smem
        if (B, T_x, T_y) in want:
            assert (K, R, max_slice, nteams) == want[(B, T_x, T_y)], (B, T_x, T_y, out.tolist())


def test_length_bucket_sampler_keeps_every_utterance_and_cuts_padding(pkg):
    """bucketing.LengthBucketBatchSampler on LJSpeech-like lengths: every index
    exactly once per epoch on every world size, different orders in different epochs, the same order for
    the same epoch, and far fewer padded cells than the reference's plain shuffle (__main__.py:237-245)."""
    from glow_tts_train_b200 import bucketing

    rng = np.random.default_rng(5)
    n = 2000
    x = rng.integers(20, 201, n)
    y = np.clip((x * 5 * rng.uniform(0.8, 1.2, n)).astype(int) // 2 * 2, x, None)
    plain = bucketing.LengthBucketBatchSampler(x, y, 32, bucket_batches=1)
    bucketed = bucketing.LengthBucketBatchSampler(x, y, 32, bucket_batches=16)
    for s in (plain, bucketed):
        batches = list(s)
        assert len(batches) == len(s) == -(-n // 32)
        assert sorted(i for b in batches for i in b) == list(range(n))
    f_plain = bucketing.padded_fraction(list(plain), x, y)
    f_bucket = bucketing.padded_fraction(list(bucketed), x, y)
    assert f_plain < 0.45 and f_bucket > 0.75 and f_bucket > 2 * f_plain, (f_plain, f_bucket)   # measured: 0.36 -> 0.80
    first = list(bucketed)
    assert first == list(bucketed)
    bucketed.set_epoch(1)
    assert first != list(bucketed)
    # two ranks: disjoint, complete, and even in alignment work batch by batch
    ranks = [bucketing.LengthBucketBatchSampler(x, y, 16, bucket_batches=16, rank=r, world_size=2) for r in range(2)]
    b0, b1 = list(ranks[0]), list(ranks[1])
    assert len(b0) == len(b1)
    assert sorted(i for b in b0 + b1 for i in b) == list(range(n))
    cost = lambda b: sum(int(x[i]) * int(y[i]) for i in b)  # noqa: E731
    worst = max(abs(cost(p) - cost(q)) / max(cost(p), cost(q)) for p, q in zip(b0[:-1], b1[:-1]))
    assert worst < 0.1, worst
    with pytest.raises(ValueError):
        bucketing.LengthBucketBatchSampler(x, y[:-1], 32)
