"""Host-side cost of one model(..., mode='sample') call at BASELINE config 2 (the GPU is idle while Python prepares the replay):
wall time per call vs GPU span, and cProfile's top functions.   python tools/call_overhead.py"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth
from subgc.config import Dims, make_opt
from subgc.model import setup

d = Dims()
m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1))
m.load_state_dict(synth.make_state_dict(d, 2019)); m.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
opt = {"beam_size": 1}
with torch.no_grad():
    for _ in range(5):
        m(*args, opt=opt, mode="sample")
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        m(*args, opt=opt, mode="sample")
    torch.cuda.synchronize()
    print(f"wall per call: {(time.perf_counter() - t0) / n * 1e3:.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        m(*args, opt=opt, mode="sample")
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)
