import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import _lib
from subgc.train import CudaOps
from subgc.config import Dims
d = Dims()
cd = _lib.Dims(d.v1, d.enc, d.rnn, d.att_hid, d.fc_feat, d.att_feat, d.gcn, d.low_rank, d.embed, d.obj_classes, d.pred_classes, d.gcn_layers, d.gcn_residual, d.pred_emb_type, d.seq_length, d.obj_num, d.rel_num)
ops = CudaOps(cd)
w = torch.randn(3000, 160, device="cuda")
xs = [torch.randn(4000, 160, device="cuda") for _ in range(64)]
out = torch.zeros(4000, 3000, device="cuda")
for mode, lst in (("same tensor", [xs[0]] * 64), ("64 distinct tensors", xs)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for x in lst:
            ops.linear(x, w, out=out, accumulate=True)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{mode}: host {1e6*(t1-t0)/64:.1f} us/call, total {1e6*(t2-t0)/64:.1f} us/call")
