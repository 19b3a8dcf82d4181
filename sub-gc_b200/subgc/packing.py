"""Split-fp16 ("h3") copies of weight matrices for the tensor-core contraction (include/subgc_b200.h: subgc_packed).

A packed copy holds the same 4 bytes per weight as the fp32 tensor (hi + lo * 2^-11, both fp16); the fp32 parameter
stays the source of truth.  `PackCache` re-packs a parameter when its storage or its version counter changed
(optimizer steps, load_state_dict), so inference after training never reads stale copies.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr


def _seg_array(cols, splits):
    pts = [0] + [int(c) for c in (splits or [])] + [int(cols)]
    assert len(pts) - 1 <= 4 and all(a < b for a, b in zip(pts, pts[1:])), pts
    return (C.c_int32 * 5)(*(pts + [0] * (5 - len(pts)))), len(pts) - 1


def pack_weight(w: torch.Tensor, splits=None, stream=None):
    """w: contiguous fp32 CUDA [rows, cols]; splits: interior column cuts of its K segments (e.g. [H, 2H] for the att-LSTM
    weight_ih).  Returns (hi, lo, overflow flag tensor, (seg array, n_seg)); hi / lo are flat int16-typed tensors in the
    k-block-major layout of subgc_packed."""
    assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.is_contiguous()
    L = lib()
    rows, cols = w.shape
    seg, n_seg = _seg_array(cols, splits)
    n = L.subgc_pack_elems(rows, n_seg, seg)
    hi = torch.empty(n, dtype=torch.int16, device=w.device)
    lo = torch.empty(n, dtype=torch.int16, device=w.device)
    flag = torch.zeros(1, dtype=torch.int32, device=w.device)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    check(L.subgc_pack_weight(rows, cols, ptr(w), cols, n_seg, seg, ptr(hi), ptr(lo), ptr(flag), st), "subgc_pack_weight")
    return hi, lo, flag, (seg, n_seg)


def packed_struct(w, hi, lo, segs):
    pk = _lib.Packed()
    pk.w, pk.hi, pk.lo = w.data_ptr(), hi.data_ptr(), lo.data_ptr()
    pk.rows, pk.cols, pk.n_seg = w.shape[0], w.shape[1], segs[1]
    for i in range(5):
        pk.seg_col[i] = segs[0][i]
    return pk


class PackCache:
    """name -> packed copy of a parameter, keyed by (data_ptr, _version)."""

    def __init__(self):
        self.entries = {}   # name -> (key, hi, lo)
        self.array = None   # ctypes array handed to the C ABI (kept alive here)
        self.array_key = None

    def build(self, named, splits=None):
        """named: dict name -> fp32 CUDA [rows, cols] parameter; splits: name -> interior column cuts of its K segments.  Returns (ctypes array, count); weights whose values do not
        fit fp16 are left out (they keep the fp32 split-TF32 path).  One host read of the overflow flags when something
        was re-packed (parameter load / update time, never inside a decode loop)."""
        fresh = []
        for n, p in named.items():
            key = (p.data_ptr(), p._version, tuple(p.shape))
            e = self.entries.get(n)
            if e is None or e[0] != key:
                hi, lo, flag, segs = pack_weight(p.detach(), (splits or {}).get(n))
                self.entries[n] = (key, hi, lo, flag, segs)
                fresh.append(n)
        for n in list(self.entries):
            if n not in named:
                del self.entries[n]
        if fresh:
            for n in fresh:
                key, hi, lo, flag, segs = self.entries[n]
                if int(flag.item()) != 0:
                    self.entries[n] = (key, None, None, None, None)
        akey = tuple((n, e[0]) for n, e in self.entries.items())
        if self.array is None or self.array_key != akey:
            good = [(n, e) for n, e in self.entries.items() if e[1] is not None]
            arr = (_lib.Packed * max(len(good), 1))()
            for i, (n, e) in enumerate(good):
                arr[i] = packed_struct(named[n], e[1], e[2], e[4])
            self.array, self.array_key, self.count = arr, akey, len(good)
        return self.array, self.count
