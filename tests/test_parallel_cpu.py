"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding reproduces the whole batch, gradient
averaging has DataParallel semantics (mean over ranks of per-rank normalised losses), variable-length row gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from subgc import parallel, synth
from subgc.config import SMALL


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(6, 3)
        dead = torch.nn.Parameter(torch.zeros(4))                    # never receives a gradient (like the dead GCN units)
        x = torch.arange(5 * 6, dtype=torch.float32).view(5, 6) / 10
        lo, hi = parallel.shard_range(5, rank, world)
        loss = lin(x[lo:hi]).pow(2).mean()                            # per-rank normalised loss
        loss.backward()
        n = parallel.allreduce_gradients(list(lin.parameters()) + [dead], world, bucket_bytes=32)
        rows = parallel.gather_rows(torch.full((rank + 1, 2), float(rank)))
        ret[rank] = (lin.weight.grad.clone(), lin.bias.grad.clone(), dead.grad, n, rows)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_has_dataparallel_semantics():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    x = torch.arange(5 * 6, dtype=torch.float32).view(5, 6) / 10
    total = 0
    for r in range(world):
        lo, hi = parallel.shard_range(5, r, world)
        total = total + lin(x[lo:hi]).pow(2).mean()
    (total / world).backward()                                        # train.py:154-156: mean over replicas
    for r in range(world):
        w, b, dead, n, rows = ret[r]
        assert torch.allclose(w, lin.weight.grad, atol=1e-6) and torch.allclose(b, lin.bias.grad, atol=1e-6)
        assert dead is None and n >= 2                                # tiny buckets -> more than one collective
        assert rows.shape == (3, 2) and rows[:, 0].tolist() == [0.0, 1.0, 1.0]


def test_shard_batch_partitions_the_loader_tensors():
    d = SMALL
    data = synth.make_train_inputs(d, 4, n_images=5, gpn_batch=2)
    for world in (1, 2, 3, 8):
        parts = [parallel.shard_batch(data, r, world) for r in range(world)]
        for k, v in data.items():
            if v is None:
                continue
            cat = torch.cat([p[k] for p in parts if p[k].shape[0] > 0])
            assert torch.equal(cat, v), k
        sizes = [p["att_feats"].shape[0] for p in parts]
        assert sum(sizes) == 5 and max(sizes) - min(sizes) <= 1
        assert all(p["labels"].shape[0] == 5 * p["att_feats"].shape[0] for p in parts)


def _reducer_worker(rank, world, port, ret):
    """Two ranks run the hand-written backward (torch emulation of the CUDA blocks) on their shard with a GradReducer; twice, so that the
    second pass uses the frozen flat-bucket layout."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_ops import EmuOps
    from helpers import load_golden, rebuild_train_case
    from subgc import train
    from subgc.model import LanguageModelCriterion
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, sd, data = rebuild_train_case(load_golden("small_train"))
        shard = parallel.shard_batch(data, rank, world)
        red = parallel.GradReducer(world)
        ops = EmuOps()
        for _ in range(2):
            with torch.no_grad():
                outputs, gpn_loss, score, S = train.forward(ops, sd, sd, d, shard, drop=None)
            leaf = outputs.clone().requires_grad_(True)
            LanguageModelCriterion()(leaf, shard["labels"][:, 1:], shard["masks"][:, 1:]).backward()
            with torch.no_grad():
                G = train.backward(ops, sd, d, S, leaf.grad, 1.0, reducer=red)
        ret[rank] = {k: v.clone() for k, v in G.items()}
    finally:
        dist.destroy_process_group()


def test_overlapped_bucket_allreduce_equals_mean_of_shard_gradients():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_ops import EmuOps
    from helpers import load_golden, rebuild_train_case
    from subgc import train
    from subgc.model import LanguageModelCriterion
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_reducer_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    d, sd, data = rebuild_train_case(load_golden("small_train"))
    ops, per = EmuOps(), []
    for r in range(world):
        shard = parallel.shard_batch(data, r, world)
        with torch.no_grad():
            outputs, gpn_loss, score, S = train.forward(ops, sd, sd, d, shard, drop=None)
        leaf = outputs.clone().requires_grad_(True)
        LanguageModelCriterion()(leaf, shard["labels"][:, 1:], shard["masks"][:, 1:]).backward()
        with torch.no_grad():
            per.append(train.backward(ops, sd, d, S, leaf.grad, 1.0))
    assert set(ret[0]) == set(per[0]) == set(ret[1])
    assert {parallel.bucket_of(n) for n in per[0]} == {0, 1, 2}
    for n in per[0]:
        mean = 0.5 * (per[0][n] + per[1][n])
        for r in range(world):
            assert torch.allclose(ret[r][n], mean, rtol=1e-5, atol=1e-7), n
