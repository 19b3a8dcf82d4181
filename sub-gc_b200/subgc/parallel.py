"""Multi-GPU plumbing for the Sub-GC path: one process per GPU, torch.distributed (NCCL over NVLink on B200, gloo in
CPU tests).

* Inference shards: images (and all their sub-graphs) are independent, so a batch is split on the image dimension and
  every rank decodes its shard — no data-path collective.  `shard_batch` performs the split on the
  loader-shaped tensors (per-image tensors on dim 0, per-sentence tensors — 5 rows per image — on dim 0 as well).
* Training is data parallel: each rank runs LossWrapper on its shard, then `allreduce_gradients` averages the gradients
  with a few large flat buckets (one NCCL all-reduce each).  Averaging per-rank *normalised* losses' gradients is exactly
  the reference's DataParallel semantics (`train.py:154-156`: mean over replicas of lang_loss / gpn_loss).
  Parameters that receive no gradient (dead GCN units, predicate embedding) are skipped consistently on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

PER_IMAGE = ("fc_feats", "att_feats", "obj_dist", "rel_ind", "pred_dist")
PER_SENTENCE = ("labels", "masks", "att_masks", "gpn_obj_ind", "gpn_pred_ind", "gpn_nrel_ind", "gpn_pool_mtx")


def shard_range(n_images: int, rank: int, world: int):
    """Contiguous image range of `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(data: dict, rank: int, world: int, seq_per_img: int = 5) -> dict:
    n_images = data["att_feats"].shape[0]
    lo, hi = shard_range(n_images, rank, world)
    out = {}
    for k, v in data.items():
        if v is None:
            out[k] = None
        elif k in PER_IMAGE:
            out[k] = v[lo:hi]
        elif k in PER_SENTENCE:
            out[k] = v[lo * seq_per_img:hi * seq_per_img]
        else:
            out[k] = v
    return out


def allreduce_gradients(params, world: int | None = None, bucket_bytes: int = 64 << 20, group=None):
    """Average .grad of `params` over the process group with flat buckets.  Returns the number of collectives issued."""
    world = world or dist.get_world_size(group)
    if world == 1:
        return 0
    live = [p for p in params if p.grad is not None]
    n_calls, bucket, size = 0, [], 0

    def flush():
        nonlocal n_calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        n_calls += 1
        bucket, size = [], 0

    for p in live:
        bucket.append(p)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_calls


def gather_rows(t: torch.Tensor, group=None):
    """Concatenate per-rank result rows (variable counts) on every rank — host-side convenience for evaluation."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device), group=group)
    mx = int(max(int(c) for c in counts))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:int(c)] for p, c in zip(parts, counts)])
