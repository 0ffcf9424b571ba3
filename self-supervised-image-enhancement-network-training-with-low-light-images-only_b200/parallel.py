"""Data-parallel exchange of the hot path: the ONLY collective is the gradient all-reduce, in two buckets.

bucket 1 = illum_adjust_net (final after phase 1 of sshslie_loss_and_grad), bucket 0 = decomposition_net (final after phase 2;
the two DecompositionNet passes share weights, model.py:231,546).  Device-agnostic on purpose: the NCCL path in
model.LowLightEnhance._dp_step and the world_size-2 gloo test (tests/test_dp_gloo.py) run the same code.
"""
import torch
import torch.distributed as dist

N_DECOMPOSITION_TENSORS = 18        # 9 layers x (weight, bias) come first in the reference's state_dict


def bucket_ranges(offsets, sizes):
    """[(lo, hi)] of the flat buffer for [decomposition_net, illum_adjust_net]."""
    n_dec = offsets[N_DECOMPOSITION_TENSORS]
    total = offsets[-1] + sizes[-1]
    return [(0, n_dec), (n_dec, total)]


def allreduce_bucket(flat, rng, group=None):
    """In-place all-reduce(sum) of one bucket (a contiguous slice, no copy)."""
    lo, hi = rng
    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)


def dp_loss_weights(weights, world):
    """Per-rank loss weights c_loss_* / world: the all-reduce(SUM) of the per-rank gradients is then the gradient of the
    global-batch loss directly (every loss term is a mean over equal shards), with no rescaling pass afterwards."""
    return [w / world for w in weights]


def finish_mean(flat, losses, world):
    """sum -> mean for a buffer that was reduced WITHOUT pre-scaled loss weights (kept for callers that exchange raw
    per-rank gradients; LowLightEnhance._dp_step uses dp_loss_weights instead)."""
    flat.mul_(1.0 / world)
    if losses is not None:
        losses.mul_(1.0 / world)
