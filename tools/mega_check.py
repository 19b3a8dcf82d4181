"""Quick GPU check of the persistent decode kernel (csrc/mega_decode.cu) against the one-launch-per-stage path: same tokens, same
log-probs, decode-stage time of both.   python tools/mega_check.py [n_images] [greedy|topk]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth
from subgc.config import Dims, make_opt
from subgc.model import setup

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "greedy"
d = Dims()
sd = synth.make_state_dict(d, 2019)
data = synth.make_test_inputs(d, 2019, n_images=n_images, per_half=1, ragged=False, ragged_edges=False)
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
res = {}
for mega in (False, True):
    m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if mode == "topk" else 0))
    m.load_state_dict(sd); m.cuda().eval()
    m.use_mega = mega
    opt = {"beam_size": 1}
    if mode == "topk":
        opt["topk_uniforms"] = torch.rand(d.seq_length, n_images, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        for _ in range(3):
            out = m(*args, opt=opt, mode="sample")
        torch.cuda.synchronize()
        m.stage_events = []
        for _ in range(10):
            out = m(*args, opt=opt, mode="sample")
        torch.cuda.synchronize()
    dec = [a.elapsed_time(b) for (n, a, b) in m.stage_events if n == "decode"]
    print(f"mega={mega}: steps={int(m.last_steps.item())} decode stage ms: min {min(dec):.3f} median {sorted(dec)[len(dec)//2]:.3f}  mega in use: {bool(m._weights().mega)}", flush=True)
    res[mega] = [t.cpu() for t in out[:2]]
same = torch.equal(res[True][0], res[False][0])
err = float((res[True][1] - res[False][1]).abs().max())
print("tokens equal:", same, " mismatches:", int((res[True][0] != res[False][0]).sum()), " max |dlogp|:", err)
print(res[True][0][:2])
print(res[False][0][:2])
sys.exit(0 if same and err < 2e-4 else 1)
