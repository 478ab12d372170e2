"""Data-parallel helpers (no counterpart in the reference, which is single-process).

Partitioning: rank r takes rows [r*B/W, (r+1)*B/W) of every global minibatch of B = W*b rows; parameters and
optimizer state are replicated; after backward the flat gradient arena is sum-allreduced and the optimizer folds
1/W.  Each rank's loss is the mean over its own b rows, so with equal shards the averaged gradient equals the
gradient of the global-batch mean loss (src/loss.rs:164).
"""
from __future__ import annotations

import numpy as np


def shard_permutation(perm, rank: int, world: int, batch_per_rank: int):
    """Per-rank index order such that step s of rank r reads rows [r*b, (r+1)*b) of global minibatch s.

    `perm` is the global sample order (DataLoader's shuffled indices, src/data/mnist.rs:326-358); the tail that
    does not fill a whole global minibatch is dropped so every rank runs the same number of steps."""
    perm = np.asarray(perm)
    gb = world * batch_per_rank
    steps = len(perm) // gb
    idx = perm[: steps * gb].reshape(steps, world, batch_per_rank)[:, rank, :]
    return np.ascontiguousarray(idx.reshape(-1)).astype(np.uint32)


def average_gradients(local_sum_grads, world: int):
    """What the step kernel does with the allreduced arena: g_mean = (sum over ranks of local mean-grads) / W."""
    return [g / np.float32(world) for g in local_sum_grads]
