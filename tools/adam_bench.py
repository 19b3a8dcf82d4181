"""HBM fraction of the fused clip + Adam step on the Sub-GC parameter set (70.05 M fp32 parameters, 280 MB): algorithmic bytes = read
p, g, m, v + write p, m, v = 7 x 280 MB per step.   python tools/adam_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc.config import Dims, make_opt
from subgc.model import setup
from subgc.optim import ClipAdam

m = setup(make_opt(Dims())).cuda()
params = [p for p in m.parameters()]
n = sum(p.numel() for p in params)
for p in params:
    p.grad = torch.randn_like(p) * 0.01
res = {}
for name, opt in (("fused", ClipAdam(params, 5e-4, clip_norm=10.0)), ("torch", torch.optim.Adam(params, 5e-4))):
    def step():
        if name == "torch":
            tot = torch.sqrt(sum(p.grad.norm(2) ** 2 for p in params))
            coef = 10.0 / max(float(tot), 10.0)                      # the reference's host round trip (misc/utils.py:196)
            for p in params:
                p.grad.mul_(coef)
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 20
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
algo = 7 * 4 * n
print(json.dumps({"params": n, "algorithmic_bytes_per_step": algo, "fused_ms": res["fused"], "torch_clip_plus_adam_ms": res["torch"],
                  "fused_gbs": algo / res["fused"] / 1e6, "hbm_peak_gbs": peak, "frac": algo / res["fused"] / 1e6 / peak,
                  "note": "the fused step reads the gradients twice (norm pass + update): its DRAM traffic is 8 x 280 MB"}))
