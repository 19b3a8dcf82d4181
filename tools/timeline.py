"""Timeline of the decode loop inside the replayed CUDA graph (SUBGC_TRACE=1): [first block start, last block end] of every launch,
printed for one token step in the middle of the loop plus per-kernel averages."""
import os, sys, ctypes as C
os.environ["SUBGC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import numpy as np
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import setup
d = Dims()
mode = sys.argv[1] if len(sys.argv) > 1 else "greedy"
model = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1))
model.load_state_dict(synth.make_state_dict(d, 2019)); model.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
L = _lib.lib()
L.subgc_debug_trace.restype = C.c_int
L.subgc_debug_trace.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
opt = {"beam_size": 5 if mode == "beam" else 1}
decode_only = mode == "head"
st0 = (C.c_ulonglong * (8 * 4096))(); ids0 = (C.c_int * 4096)()
with torch.no_grad():
    model(*args, opt=opt, mode="sample")            # eager
    L.subgc_debug_trace(0, None, None, 0)           # slots restart: the capture call numbers the graph's launches 0..n
    model(*args, opt=opt, mode="sample")            # capture + replay
    n_slots = L.subgc_debug_trace(2, st0, ids0, 4096)
    for _ in range(3):
        model(*args, opt=opt, mode="sample")
    L.subgc_debug_trace(1, None, None, 0)
    model(*args, opt=opt, mode="sample")
torch.cuda.synchronize()
N = 4096
st = (C.c_ulonglong * (8 * N))(); ids = (C.c_int * N)()
# slot numbering was restarted before the capture call, but eager launches of later calls keep counting: read the first slots
m = L.subgc_debug_trace(2, st, ids, N)
t = np.array(st, dtype=np.float64).reshape(N, 8)[:m]
k = np.array(ids[:m])
ok = (t[:, 1] > 0) & (t[:, 0] < 1e19)
print("slots after capture call:", n_slots, "slots now:", m, "ids histogram:", {int(i): int((k == i).sum()) for i in np.unique(k)},
      "valid per id:", {int(i): int((ok & (k == i)).sum()) for i in np.unique(k)})
names = {1: "h3_gemm", 2: "cell", 3: "attention", 4: "select", 5: "att_phase", 6: "h3_gemm+cell"}
t0 = t[ok, 0].min()
idx = np.nonzero(ok)[0]
sel = [i for i in idx if k[i] == 4]
print(f"{ok.sum()} traced launches, loop span {(t[ok,1].max()-t0)/1e3:.1f} us, {len(sel)} select launches")
if mode == "head":
    # the start of the loop: everything before the third selection
    z = t[idx[0], 0]
    for i in idx[:22]:
        print(f"  {names.get(int(k[i]), str(k[i])):12s} start {(t[i,0]-z)/1e3:8.2f} released {(t[i,2]-z)/1e3:8.2f} end {(t[i,1]-z)/1e3:8.2f}")
    print(f"  last kernel end {(t[idx[-1],1]-z)/1e3:8.2f}")
if mode == "beam":
    # no selection kernel in the beam loop: print 16 consecutive launches from the middle of the loop
    mid = idx[len(idx) // 2: len(idx) // 2 + 16]
    z = t[mid[0], 0]
    for i in mid:
        print(f"  {names.get(int(k[i]), str(k[i])):12s} start {(t[i,0]-z)/1e3:8.2f} released {(t[i,2]-z)/1e3:8.2f} end {(t[i,1]-z)/1e3:8.2f}  dur {(t[i,1]-t[i,0])/1e3:7.2f}")
if len(sel) >= 12:
    a, b = sel[9] + 1, sel[10] + 1
    prev_end = t[sel[9], 1]
    print("one step (us relative to the previous select's end):  start  released(first..last)    end   | end - prev end")
    for i in range(a, b):
        if not ok[i]:
            continue
        z = t[sel[9], 1]
        print(f"  {names.get(int(k[i]), str(k[i])):12s} {(t[i,0]-z)/1e3:8.2f} {(t[i,2]-z)/1e3:8.2f} ..{(t[i,3]-z)/1e3:7.2f} {(t[i,1]-z)/1e3:8.2f}   | {(t[i,1]-prev_end)/1e3:7.2f}   marks " + " ".join(f"{(t[i,4+j]-z)/1e3:7.2f}" if t[i,4+j] > 0 else "   -   " for j in range(4)))
        prev_end = t[i, 1]
    print(f"  step total {(t[sel[10],1]-t[sel[9],1])/1e3:.2f} us")
for kid, nm in names.items():
    m_ = ok & (k == kid)
    if m_.any():
        print(f"{nm:10s} n={m_.sum():4d} mean dur {np.mean(t[m_,1]-t[m_,0])/1e3:7.2f} us")
