"""Timeline of the persistent decode kernel (SUBGC_MEGA_TRACE=1): per step, when each hand-off happens (globaltimer stamps written by
worker thread 0 / the MMA thread / the producer of every CTA).   python tools/mega_timeline.py [n_images] [step]"""
import os, sys, ctypes as C
os.environ["SUBGC_MEGA_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import numpy as np
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import setup

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 128
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = Dims()
m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if os.environ.get("MG_TOPK") else 0))
m.load_state_dict(synth.make_state_dict(d, 2019)); m.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=n_images, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
with torch.no_grad():
    for _ in range(5):
        m(*args, opt={"beam_size": 1}, mode="sample")
torch.cuda.synchronize()
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
T, EV = d.seq_length, 48
buf = (C.c_ulonglong * (n_cta * T * EV))()
_lib.check(_lib.lib().subgc_debug_mega_trace(buf, n_cta, T), "trace")
a = np.array(buf, dtype=np.float64).reshape(n_cta, T, EV)
a[a == 0] = np.nan
names = {0: "step begins (workers)", 1: "acc A(t+1) complete", 2: "partial A(t+1) published", 3: "tokens + tile A partials seen (cell A)", 4: "h_att published",
         5: "acc B complete", 6: "partial B published", 7: "all B partials seen (attention)", 8: "ctx published", 9: "acc C complete",
         10: "partial C published", 11: "tile C partials complete", 12: "h_lang published", 13: "acc D complete", 14: "partial D published",
         15: "all D partials seen (select)", 16: "token published", 40: "producer: step issued",
         36: "  cell A operands landed", 37: "  cell C operands landed", 17: "  cell A computed + stored", 19: "  cell C computed + stored", 41: "  atth reduced", 42: "  scores done", 43: "  softmax done",
         38: "  top-k: scaled log-probs + exp done", 39: "  top-k: sum and candidates merged", 44: "  ctx partials done", 18: "  ctx stored", 45: "  logits loaded", 46: "  argmax + lse done", 47: "  token chosen"}
tasks = ["C_hlang", "B", "C_hatt", "A_hatt+", "C_ctx", "D", "A_hlang+"]
t0 = np.nanmin(a[:, step, 0])
print(f"step {step} of {T}; us relative to the first CTA entering the step; min / median / max over CTAs")
rows = []
for ev in sorted(names):
    v = (a[:, step, ev] - t0) / 1e3
    if np.all(np.isnan(v)):
        continue
    rows.append((np.nanmedian(v), f"  {names[ev]:36s} {np.nanmin(v):8.2f} {np.nanmedian(v):8.2f} {np.nanmax(v):8.2f}   (n={int(np.sum(~np.isnan(v)))})"))
# MMA tasks: the task index differs between CTAs that have / have not a B task; report by CTA class
for cls, sel in (("CTAs without B", np.isnan(a[:, step, 5])), ("CTAs with B", ~np.isnan(a[:, step, 5]))):
    if not sel.any():
        continue
    for k in range(8):
        v0 = (a[sel, step, 20 + 2 * k] - t0) / 1e3
        v1 = (a[sel, step, 21 + 2 * k] - t0) / 1e3
        if np.all(np.isnan(v0)):
            continue
        rows.append((np.nanmedian(v0), f"  MMA task {k} [{cls}] operands landed   {np.nanmin(v0):8.2f} {np.nanmedian(v0):8.2f} {np.nanmax(v0):8.2f}"))
        rows.append((np.nanmedian(v1), f"  MMA task {k} [{cls}] issued            {np.nanmin(v1):8.2f} {np.nanmedian(v1):8.2f} {np.nanmax(v1):8.2f}"))
for _, line in sorted(rows, key=lambda r: r[0]):
    print(line)
per_step = (np.nanmax(a[:, 1:, 16], axis=0) - np.nanmax(a[:, :-1, 16], axis=0)) / 1e3
print("step time (token published -> next token published), us:", np.array2string(per_step, precision=1))
print("kernel span us:", (np.nanmax(a[:, T - 1, 16]) - np.nanmin(a[:, 0, 0])) / 1e3)
if os.environ.get("MG_OUTLIERS"):
    # which CTAs are late, per step: the hand-offs wait for the slowest one
    for ev in [int(x) for x in os.environ.get('MG_OUTLIER_EVENTS', '2,16,14,10,4,12').split(',')]:
        print(f"late CTAs for '{names[ev].strip()}' (step: cta +us over the median)")
        for s in range(T):
            v = (a[:, s, ev] - np.nanmedian(a[:, s, ev])) / 1e3
            if np.all(np.isnan(v)):
                continue
            order = np.argsort(np.nan_to_num(v, nan=-1e9))[::-1][:3]
            print(f"  {s:2d}: " + "  ".join(f"{int(c)} +{v[c]:.1f}" for c in order))
