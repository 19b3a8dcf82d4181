import os, sys
os.environ["SUBGC_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import _lib
L = _lib.lib()
M, N, K = (int(v) for v in os.environ.get('GEMM_TRACE_SHAPE', '128,4000,4000').split(','))
A = torch.randn(M, K).cuda(); W = (torch.randn(N, K) / 64).cuda(); b = torch.zeros(N).cuda()
out = torch.empty(M, N, device="cuda")
ws = torch.empty(L.subgc_linear_workspace_bytes(M, N, K) + 256, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    flush.zero_(); torch.cuda.synchronize()
    _lib.check(L.subgc_linear_forward(M, N, K, A.data_ptr(), K, None, W.data_ptr(), K, (None if os.environ.get('GEMM_TRACE_NOBIAS') else b.data_ptr()), 0, out.data_ptr(), N, ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream), "linear")
    torch.cuda.synchronize()
