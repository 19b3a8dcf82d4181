import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import synth, _lib
from subgc.config import Dims, make_opt
from subgc.model import LossWrapper, setup
d = Dims()
sd = synth.make_state_dict(d, 1)
model = setup(make_opt(d)); model.load_state_dict(sd); model.cuda().train()
lw = LossWrapper(model, None)
data = synth.make_train_inputs(d, 1, n_images=32, gpn_batch=2)
data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
call = (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
        data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
L = _lib.lib()
for it in range(4):
    for p in model.parameters(): p.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = L.subgc_launch_count()
    out = lw(*call)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    (out["lang_loss"] + out["gpn_loss"]).backward()
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"iter {it}: fwd host {1e3*(t1-t0):.1f} ms (+sync {1e3*(t2-t1):.1f}), bwd host {1e3*(t3-t2):.1f} ms (+sync {1e3*(t4-t3):.1f}), launches {L.subgc_launch_count()-n0}", flush=True)
