"""Kernel-by-kernel GPU timeline of ONE replay of the whole-call CUDA graph (encoder .. persistent decode kernel), from CUPTI through
torch.profiler: start offset, duration and gap to the previous kernel.   python tools/graph_kernels.py [greedy|topk|beam]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from subgc import synth
from subgc.config import Dims, make_opt
from subgc.model import setup

mode = sys.argv[1] if len(sys.argv) > 1 else "greedy"
d = Dims()
m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if mode == "topk" else 0))
m.load_state_dict(synth.make_state_dict(d, 2019)); m.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
opt = {"beam_size": 5 if mode == "beam" else 1}
with torch.no_grad():
    for _ in range(5):
        m(*args, opt=opt, mode="sample")
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(*args, opt=opt, mode="sample")
        torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0, prev_end = ev[0].time_range.start, ev[0].time_range.start
print(f"{'start us':>9s} {'dur us':>8s} {'gap us':>7s}  kernel")
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    print(f"{s - t0:9.1f} {t - s:8.1f} {s - prev_end:7.1f}  {e.name[:90]}")
    prev_end = max(prev_end, t)
print(f"span {prev_end - t0:.1f} us, {len(ev)} kernels, busy {sum(e.time_range.end - e.time_range.start for e in ev):.1f} us")
