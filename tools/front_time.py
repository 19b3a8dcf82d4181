"""GPU time of the front-end stages (encoder, sGPN + NMS, prepare) when each replays as its own CUDA graph, interleaved with the
decode graph so that caches are in the state of a real step.  python tools/front_time.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import setup, _capture_stream

d = Dims()
m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1))
m.load_state_dict(synth.make_state_dict(d, 2019)); m.cuda().eval()
data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
with torch.no_grad():
    for _ in range(3):
        m(*args, opt={"beam_size": 1}, mode="sample")
    torch.cuda.synchronize()
    state = {}

    def s_encode():
        state["x_obj"] = m.encode(dev["att_feats"], dev["obj_dist"], dev["pred_dist"], dev["rel_ind"])

    def s_sgpn():
        state["sg"] = m._sgpn(state["x_obj"], dev["gpn_obj_ind"], dev["att_masks"], order=1)

    def s_prepare():
        lay, n_sub, read_out, score, sub_len, loss = state["sg"]
        sel = torch.arange(0, 256, 2, dtype=torch.int32, device="cuda")
        state["prep"] = m._prepare(lay, 128, 37, sel, state["x_obj"], dev["gpn_obj_ind"], dev["att_masks"], read_out)

    graphs = []
    for name, fn in (("encode", s_encode), ("sgpn", s_sgpn), ("prepare", s_prepare)):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        c0 = _lib.lib().subgc_launch_count()
        with torch.cuda.graph(g, stream=_capture_stream(torch.device("cuda", 0))):
            fn()
        graphs.append((name, g, int(_lib.lib().subgc_launch_count() - c0)))
    tot = {n: 0.0 for n, _, _ in graphs}
    R = 20
    for it in range(R + 3):
        evs = []
        for name, g, _ in graphs:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            evs.append((name, a, b))
        m(*args, opt={"beam_size": 1}, mode="sample")   # the rest of the step (thrashes L2 like the real thing)
        torch.cuda.synchronize()
        if it >= 3:
            for name, a, b in evs:
                tot[name] += a.elapsed_time(b)
    for name, g, n in graphs:
        print(f"{name:8s} {tot[name] / R * 1e3:8.1f} us per replay, {n} kernels")
