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


# Reverse-autograd order of the hand-written backward (subgc.train.backward): the decoder's parameters are final after the BPTT loop,
# then feature preparation, then sGPN / GCN / fusion ("logit/embed -> LSTMs -> prepare -> sGPN -> GCN").
BUCKET_OF_PREFIX = (("logit.", 0), ("embed.", 0), ("core.", 0), ("fc_embed.", 1), ("att_embed.", 1), ("ctx2att.", 1),
                    ("gpn_layer.read_out_proj.", 1), ("gpn_layer.gpn_fc.", 2), ("gcn_backbone.", 2), ("obj_v_proj.", 2), ("obj_emb_proj.", 2),
                    ("sg_obj_embed.", 2), ("sg_pred_embed.", 2), ("pred_emb_prj.", 2))
N_BUCKETS = 3


def bucket_of(name: str) -> int:
    for prefix, b in BUCKET_OF_PREFIX:
        if name.startswith(prefix):
            return b
    return N_BUCKETS - 1


class GradReducer:
    """Gradient all-reduce overlapped with the backward pass.  The hand-written backward allocates every gradient as a
    VIEW into one flat buffer per bucket (`alloc`), and calls `bucket_done(b)` as soon as the last kernel writing into bucket b has been
    enqueued: the collective of that bucket (one NCCL all-reduce over the flat buffer, no flatten / copy-back) then runs on the
    communicator's stream while the remaining backward kernels run on the compute stream.  `finish()` makes the compute stream wait for
    the collectives and applies the 1 / world factor: mean over ranks of per-rank normalised losses = the reference's DataParallel
    semantics (train.py:154-156)."""

    def __init__(self, world: int | None = None, group=None):
        self.group = group
        self.world = world or (dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1)
        self.sizes = None        # per bucket: {name: (offset, numel)} fixed by the first backward pass
        self.flat = [None] * N_BUCKETS
        self.work = [None] * N_BUCKETS
        self.used = [0] * N_BUCKETS
        self.plan = [dict() for _ in range(N_BUCKETS)]
        self.frozen = False      # after the first pass the layout is fixed and the flat buffers are re-used
        self.exposed_ms = None

    def begin(self):
        for b in range(N_BUCKETS):
            self.work[b] = None
            if self.frozen and self.flat[b] is not None:
                # a fresh buffer per step: the gradients handed to autograd last step are views of the old one and stay valid
                self.flat[b] = torch.zeros_like(self.flat[b])

    def alloc(self, name: str, like: torch.Tensor) -> torch.Tensor:
        b = bucket_of(name)
        n = like.numel()
        if self.frozen:
            off, m = self.plan[b][name]
            assert m == n
            return self.flat[b][off:off + n].view_as(like)
        # first pass: plain tensors; the flat layout is built from them when the bucket closes
        t = torch.zeros_like(like)
        self.plan[b][name] = t
        return t

    def bucket_done(self, b: int, G: dict):
        names = [n for n in G if bucket_of(n) == b]
        if not names:
            return
        if not self.frozen:
            # build the flat buffer of this bucket from the tensors of the first pass (one extra copy, once)
            flat = torch.cat([G[n].reshape(-1) for n in names])
            off, plan = 0, {}
            for n in names:
                m = G[n].numel()
                plan[n] = (off, m)
                G[n] = flat[off:off + m].view_as(G[n])
                off += m
            self.flat[b], self.plan[b] = flat, plan
        if self.world > 1:
            self.work[b] = dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self, G: dict):
        ev0 = ev1 = None
        cuda = any(f is not None and f.is_cuda for f in self.flat)
        if cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for b in range(N_BUCKETS):
            if self.work[b] is not None:
                self.work[b].wait()           # the compute stream waits for the collective; the host does not block on CUDA
                self.flat[b].div_(self.world)
        if cuda:
            ev1.record()
            self._exposed = (ev0, ev1)       # time the compute stream spent waiting for / scaling the collectives
        self.frozen = True
        return G

    def exposed_wait_ms(self):
        """GPU time between the end of the backward kernels and the end of the last collective (+ the 1/world scaling)."""
        if getattr(self, "_exposed", None) is None:
            return None
        a, b = self._exposed
        b.synchronize()
        return a.elapsed_time(b)


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
