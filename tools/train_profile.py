"""Where a training step (BASELINE config 5 shard: LossWrapper forward + backward) spends its time: wall clock per step, GPU kernel time by
kernel name (torch.profiler / CUPTI), launches per step.   python tools/train_profile.py [n_images]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import LossWrapper, setup

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = Dims()
dev = torch.device("cuda", 0)
model = setup(make_opt(d)); model.load_state_dict(synth.make_state_dict(d, 2019)); model.to(dev).train()
lw = LossWrapper(model, None)
data = synth.make_train_inputs(d, 2019, n_images=n_images, gpn_batch=2, ragged=True, ragged_edges=True)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
call = (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
        data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
params = list(model.parameters())


def step():
    for p in params:
        p.grad = None
    out = lw(*call)
    (out["lang_loss"] + out["gpn_loss"]).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
# host-side time of the stage-level C calls and of the two halves of a step (no synchronisation: issue time only)
from subgc import train as _train
acc = {}


def _timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w


for nm in ("decoder_forward", "decoder_backward", "linear", "fuse_nodes", "transpose", "colsum"):
    if hasattr(_train.CudaOps, nm):
        setattr(_train.CudaOps, nm, _timed(nm, getattr(_train.CudaOps, nm)))
for p_ in params:
    p_.grad = None
t = time.perf_counter(); out = lw(*call); t_f = time.perf_counter() - t
t = time.perf_counter(); (out["lang_loss"] + out["gpn_loss"]).backward(); t_b = time.perf_counter() - t
torch.cuda.synchronize()
print(f"host issue: forward {t_f * 1e3:.2f} ms, backward {t_b * 1e3:.2f} ms; inside: " + ", ".join(f"{k} {v * 1e3:.2f}" for k, v in acc.items()))
L = _lib.lib()
c0 = L.subgc_launch_count()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t1 = time.perf_counter() - t0
print(f"wall per step {t1 / 5 * 1e3:.2f} ms (host issue {t_issue / 5 * 1e3:.2f} ms), subgc launches per step {(L.subgc_launch_count() - c0) / 5:.0f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key) for e in ka if (getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)) > 0], reverse=True)
tot = sum(r[0] for r in rows)
print(f"GPU kernel time of one step: {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for us, n, k in rows[:40]:
    print(f"  {us / 1e3:8.3f} ms {n:5d}x  {k[:110]}")

print("host side (self CPU time):")
for e in sorted(ka, key=lambda e: -e.self_cpu_time_total)[:25]:
    print(f"  {e.self_cpu_time_total / 1e3:8.3f} ms {e.count:5d}x  {e.key[:100]}")
