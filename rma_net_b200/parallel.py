"""Multi-GPU plumbing for the chamfer path (one process per GPU, torch.distributed).

* Batched configs (C2-C4): batch elements never interact (loss.py:51-57 is a per-b bmm, the loss is a sum over b), so
  the batch is sharded contiguously and NO collective touches the data path; `shard_range` / `shard_batch` are all
  there is to it (an optional scalar all-reduce gives the global loss).
* One huge pair (C5): the SOURCE cloud is sharded, the target replicated.  source->target is rank-local and complete.
  target->source: every rank finds, for all M targets, the best of ITS sources; the winners are merged with ONE
  all-reduce-MIN over packed keys  float_bits(d) << 32 | global_source_index  (d >= 0, so integer order is
  (distance, index) order and ties go to the lowest index, torch.min semantics).  Backward: grad_source is local
  once the winners are known (the owner rank of idx2[j] applies that term); grad_target partial sums are merged
  with one all-reduce-SUM.

The reference has no distributed code at all; this is new functionality, not a port.
The compute callables default to the CUDA ops; tests inject CPU stand-ins to exercise this logic under gloo.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) of n items for `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's contiguous slice of the batch dimension (a view; the kernels take strided inputs)."""
    s, e = shard_range(t.shape[0], rank, world)
    return t[s:e]


def global_loss(local_loss_sum: torch.Tensor, local_batch: int, group=None) -> torch.Tensor:
    """Mean-over-global-batch loss from per-rank SUMS over local batch elements (loss_on_cd divides by B, loss.py:205).
    The only collective of the batched mode, and it is off the data path."""
    v = torch.stack([local_loss_sum.detach().to(torch.float64).reshape(()),
                     torch.tensor(float(local_batch), dtype=torch.float64, device=local_loss_sum.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return (v[0] / v[1]).to(torch.float32)


class ShardOps(NamedTuple):
    """Compute callables used by the source-sharded mode (CUDA ops by default)."""
    chamfer_forward: Callable      # (x [B,n,3], y [B,M,3]) -> dist1, dist2, idx1, idx2 (int32)
    chamfer_backward: Callable     # (x, y, idx1, idx2, g1, g2, need_x, need_y) -> grad_x, grad_y
    pack_min_keys: Callable        # (dist f32, idx i32, offset) -> int64 keys
    unpack_min_keys: Callable      # keys -> (dist f32, idx i32)


def cuda_ops() -> ShardOps:
    from . import ext
    return ShardOps(ext.chamfer_forward, ext.chamfer_backward, ext.pack_min_keys, ext.unpack_min_keys)


def merge_min_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce-MIN of packed (distance, index) keys (int64; sign bit never set because d >= 0)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def owned_winners(idx2_global: torch.Tensor, offset: int, n_local: int):
    """Which targets were won by a source of this rank, and that source's local index (0 where not owned)."""
    owned = (idx2_global >= offset) & (idx2_global < offset + n_local)
    local = torch.where(owned, idx2_global - offset, torch.zeros_like(idx2_global))
    return owned, local.to(torch.int32)


def source_sharded_forward(x_local: torch.Tensor, y: torch.Tensor, offset: int, group=None,
                           ops: Optional[ShardOps] = None):
    """x_local [B,n_local,3] = sources [offset, offset+n_local) of the global cloud; y [B,M,3] replicated.
    Returns dist1 [B,n_local], dist2 [B,M] (global), idx1 [B,n_local] (target index), idx2 [B,M] (GLOBAL source)."""
    ops = ops or cuda_ops()
    d1, d2_local, i1, i2_local = ops.chamfer_forward(x_local, y)
    keys = ops.pack_min_keys(d2_local, i2_local, offset)
    merge_min_keys(keys, group)
    d2, i2 = ops.unpack_min_keys(keys)
    return d1, d2, i1, i2


def source_sharded_backward(x_local, y, offset, idx1, idx2_global, g1, g2, group=None,
                            ops: Optional[ShardOps] = None, need_x: bool = True, need_y: bool = True):
    """Gradients of the source-sharded chamfer: grad_x_local [B,n_local,3] (complete), grad_y [B,M,3] (all-reduced)."""
    ops = ops or cuda_ops()
    owned, idx2_local = owned_winners(idx2_global, offset, x_local.shape[1])
    g2_owned = torch.where(owned, g2, torch.zeros_like(g2)) if g2 is not None else None
    gx, gy = ops.chamfer_backward(x_local, y, idx1, idx2_local, g1, g2_owned, need_x, need_y)
    if gy is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gy, op=dist.ReduceOp.SUM, group=group)
    return gx, gy


class SourceShardedChamfer(torch.autograd.Function):
    """chamfer_dist for ONE huge pair with the source cloud sharded over the ranks of `group`."""

    @staticmethod
    def forward(ctx, x_local, y, offset, group):
        d1, d2, i1, i2 = source_sharded_forward(x_local, y, offset, group)
        ctx.save_for_backward(x_local, y, i1, i2)
        ctx.offset, ctx.group = offset, group
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, g1, g2, _a, _b):
        x_local, y, i1, i2 = ctx.saved_tensors
        gx, gy = source_sharded_backward(x_local, y, ctx.offset, i1, i2, g1, g2, ctx.group,
                                         need_x=ctx.needs_input_grad[0], need_y=ctx.needs_input_grad[1])
        return gx, gy, None, None


def chamfer_dist_source_sharded(x_local, y, offset: int, group=None):
    """-> (dist1 [B,n_local], dist2 [B,M]); differentiable.  Every rank must call it (collectives inside)."""
    d1, d2, _, _ = SourceShardedChamfer.apply(x_local, y, offset, group)
    return d1, d2
