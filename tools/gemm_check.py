"""Accuracy / timing probe of the contraction block through the C ABI: error vs fp64 for the split-fp16 tensor-core path
(packed weights, subgc_linear_packed_forward), the split-TF32 path (subgc_linear_forward) and, with SUBGC_GEMM=simt, the
fp32 FMA path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
import torch
from subgc import _lib, packing
import ctypes as C

L = _lib.lib()
print("mode:", os.environ.get("SUBGC_GEMM", "tc"))
shapes = [(128, 256, 64), (128, 512, 1000), (128, 4000, 4000), (128, 4000, 3000), (37, 64, 48), (130, 9488, 1000), (4736, 1024, 2048), (5, 4000, 3000)]
for (M, N, K) in shapes:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(A.double(), W.double(), b.double())
    ref32 = torch.nn.functional.linear(A, W, b)
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    ws = torch.empty(L.subgc_linear_workspace_bytes(M, N, K) + 256, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    hi, lo, flag, segs = packing.pack_weight(Wd)
    pk = packing.packed_struct(Wd, hi, lo, segs)
    scale = ref.abs().max()
    line = f"M={M} N={N} K={K}: (torch-cpu-fp32 {float((ref32.double()-ref).abs().max()/scale):.2e})"
    for name in ("h3", "tf32x3"):
        out = torch.full((M, N), float("nan"), device="cuda")
        def run():
            if name == "h3":
                _lib.check(L.subgc_linear_packed_forward(M, N, K, Ad.data_ptr(), K, None, C.byref(pk), bd.data_ptr(), 0, out.data_ptr(), N,
                                                         ws.data_ptr(), ws.numel(), st), "linear_packed")
            else:
                _lib.check(L.subgc_linear_forward(M, N, K, Ad.data_ptr(), K, None, Wd.data_ptr(), K, bd.data_ptr(), 0, out.data_ptr(), N,
                                                  ws.data_ptr(), ws.numel(), st), "linear")
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        o = out.cpu().double()
        line += f"  | {name}: err {float((o-ref).abs().max()/scale):.2e} nan={int(torch.isnan(o).sum())} {e0.elapsed_time(e1)/10*1e3:.1f} us"
    print(line + f"  ovf={int(flag.item())}", flush=True)
