"""Batch shaping.  The reference pads every batch to its longest utterance
(PhonemeMelCollate, dataset.py:77-116) and draws batches from a plain shuffle (__main__.py:237-245),
so the padded cell count B * max(t_x) * max(t_y) -- what the alignment path, the encoder and the
decoder all pay for -- is typically 1.7x the valid one.  ``LengthBucketBatchSampler`` is a
``batch_sampler`` for ``torch.utils.data.DataLoader`` that keeps the shuffle between epochs but puts
utterances of similar length into the same batch; with ``world_size > 1`` every global batch is dealt
to the ranks by ``sharding.balanced_shards`` so that each rank's share of the alignment work is even.
Pure host logic (no torch needed to construct it)."""
from __future__ import annotations

import random
from typing import Iterator, List, Sequence

from . import sharding


def padded_fraction(batches: Sequence[Sequence[int]], x_lengths: Sequence[int], y_lengths: Sequence[int]) -> float:
    """valid cells / padded cells over `batches` (cells = tokens x frames, the alignment lattice)."""
    valid = padded = 0
    for b in batches:
        if not b:
            continue
        valid += sum(x_lengths[i] * y_lengths[i] for i in b)
        padded += len(b) * max(x_lengths[i] for i in b) * max(y_lengths[i] for i in b)
    return valid / padded if padded else 1.0


class LengthBucketBatchSampler:
    """Yields lists of dataset indices.  Every epoch: shuffle all indices (seed + epoch), cut the
    shuffled order into windows of ``bucket_batches`` global batches, sort each window by frame count
    (descending, ties by token count) and cut it into global batches; the ORDER of the batches is then
    shuffled again.  Randomness between epochs is kept at window granularity, padding shrinks to what a
    window's length spread allows (bucket_batches = 1 is the reference's plain shuffle).
    ``rank`` / ``world_size``: this rank's part of each global batch, balanced by t_x * t_y."""

    def __init__(self, x_lengths: Sequence[int], y_lengths: Sequence[int], batch_size: int, *, bucket_batches: int = 16,
                 seed: int = 1234, rank: int = 0, world_size: int = 1, drop_last: bool = False):
        if len(x_lengths) != len(y_lengths):
            raise ValueError("x_lengths and y_lengths must describe the same utterances")
        if batch_size < 1 or bucket_batches < 1 or world_size < 1 or not 0 <= rank < world_size:
            raise ValueError("bad batch_size / bucket_batches / rank / world_size")
        self.x, self.y = list(map(int, x_lengths)), list(map(int, y_lengths))
        self.batch_size, self.bucket_batches = batch_size, bucket_batches
        self.seed, self.rank, self.world_size, self.drop_last = seed, rank, world_size, drop_last
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def global_batches(self) -> List[List[int]]:
        rng = random.Random(self.seed * 1000003 + self.epoch)
        order = list(range(len(self.x)))
        rng.shuffle(order)
        gb = self.batch_size * self.world_size
        window = gb * self.bucket_batches
        batches: List[List[int]] = []
        for w in range(0, len(order), window):
            chunk = sorted(order[w:w + window], key=lambda i: (-self.y[i], -self.x[i], i))
            for s in range(0, len(chunk), gb):
                b = chunk[s:s + gb]
                if len(b) == gb or not self.drop_last:
                    batches.append(b)
        rng.shuffle(batches)
        return batches

    def __iter__(self) -> Iterator[List[int]]:
        for b in self.global_batches():
            if self.world_size == 1:
                yield b
                continue
            shards = sharding.balanced_shards([self.x[i] * self.y[i] for i in b], self.world_size)
            yield [b[j] for j in shards[self.rank]]

    def __len__(self) -> int:
        gb = self.batch_size * self.world_size
        n = len(self.x)
        return n // gb if self.drop_last else -(-n // gb)
