"""Per-stage timing of the fused att-phase kernel inside the replayed decode graph (SUBGC_ATT_TRACE=1): globaltimer stamps of the
last launch, per block: start | cell done | barrier 1 passed | h2att done | barrier 2 passed | attention done."""
import os, sys, ctypes as C
os.environ["SUBGC_ATT_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import setup
d = Dims()
model = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1))
model.load_state_dict(synth.make_state_dict(d, 2019)); model.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
with torch.no_grad():
    for _ in range(5):
        model(*args, opt={"beam_size": 1}, mode="sample")
torch.cuda.synchronize()
L = _lib.lib()
L.subgc_debug_att_trace.restype = C.c_int
buf = (C.c_ulonglong * (128 * 8))()
assert L.subgc_debug_att_trace(buf, 128) == 0
import numpy as np
t = np.array(buf, dtype=np.float64).reshape(128, 8)[:, :6]
t0 = t[:, 0].min()
rel = t - t0
names = ["start", "cell", "bar1", "h2att", "bar2", "attn"]
print("ns since first block start: mean / max per stamp")
for i, n in enumerate(names):
    print(f"  {n:6s} mean {rel[:, i].mean():8.0f}  min {rel[:, i].min():8.0f}  max {rel[:, i].max():8.0f}")
