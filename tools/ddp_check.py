"""Run under torchrun on 2+ GPUs: data-parallel gradients (per-rank LossWrapper backward + NCCL all-reduce) must equal the
mean over shards computed by one process (DataParallel semantics of the reference, train.py:154-156)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
import torch.distributed as dist
from subgc import parallel, synth
from subgc.config import SMALL, make_opt
from subgc.model import LossWrapper, setup

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = SMALL
sd = synth.make_state_dict(d, 7, logit_gain=4.0)
full = synth.make_train_inputs(d, 7, n_images=2 * world, gpn_batch=2)
model = setup(make_opt(d)); model.load_state_dict(sd); model.to(dev).train(); model.dropout_enabled = False
lw = LossWrapper(model, None)

def run(data):
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    for p in model.parameters():
        p.grad = None
    o = lw(data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
           data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
    (o["lang_loss"] + o["gpn_loss"]).backward()
    return {n: (p.grad.clone() if p.grad is not None else None) for n, p in model.named_parameters()}

# (a) plain bucketed all-reduce after backward, (b) the reducer that overlaps the three bucket collectives with the backward (twice: the
# second pass runs on the frozen flat-bucket layout)
mine = run(parallel.shard_batch(full, rank, world))
for n, p in model.named_parameters():
    p.grad = mine[n]
calls = parallel.allreduce_gradients(list(model.parameters()), world)
got = {n: (p.grad.clone() if p.grad is not None else None) for n, p in model.named_parameters()}
model.grad_reducer = parallel.GradReducer(world)
for _ in range(2):
    got_overlapped = run(parallel.shard_batch(full, rank, world))
exposed = model.grad_reducer.exposed_wait_ms()
model.grad_reducer = None
ref = None
for r in range(world):
    g = run(parallel.shard_batch(full, r, world))
    ref = g if ref is None else {n: (None if v is None else v + g[n]) for n, v in ref.items()}
worst = 0.0
for n, v in ref.items():
    if v is None:
        assert got[n] is None, n
        continue
    v = v / world
    err = float((got[n] - v).abs().max() / (v.abs().max() + 1e-12))
    err2 = float((got_overlapped[n] - v).abs().max() / (v.abs().max() + 1e-12))
    worst = max(worst, err, err2)
assert worst < 1e-5, worst
if rank == 0:
    print(f"ddp_check ok: world={world}, {calls} all-reduce calls (plain) / 3 overlapped bucket collectives, worst relative gradient error "
          f"{worst:.2e}, exposed wait {exposed:.3f} ms")
dist.destroy_process_group()
