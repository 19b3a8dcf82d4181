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


def pack_weight(w: torch.Tensor, stream=None):
    """w: contiguous fp32 CUDA [rows, cols].  Returns (hi, lo, overflowed) with hi/lo int16-typed [rows, ld16] tensors."""
    assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.is_contiguous()
    L = lib()
    rows, cols = w.shape
    ld16 = L.subgc_pack_ld(cols)
    hi = torch.empty(rows, ld16, dtype=torch.int16, device=w.device)
    lo = torch.empty(rows, ld16, dtype=torch.int16, device=w.device)
    flag = torch.zeros(1, dtype=torch.int32, device=w.device)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    check(L.subgc_pack_weight(rows, cols, ptr(w), cols, ptr(hi), ptr(lo), ptr(flag), st), "subgc_pack_weight")
    return hi, lo, flag


def packed_struct(w, hi, lo):
    pk = _lib.Packed()
    pk.w, pk.hi, pk.lo = w.data_ptr(), hi.data_ptr(), lo.data_ptr()
    pk.rows, pk.cols, pk.ld16 = w.shape[0], w.shape[1], hi.shape[1]
    return pk


class PackCache:
    """name -> packed copy of a parameter, keyed by (data_ptr, _version)."""

    def __init__(self):
        self.entries = {}   # name -> (key, hi, lo)
        self.array = None   # ctypes array handed to the C ABI (kept alive here)
        self.array_key = None

    def build(self, named):
        """named: dict name -> fp32 CUDA [rows, cols] parameter.  Returns (ctypes array, count); weights whose values do not
        fit fp16 are left out (they keep the fp32 split-TF32 path).  One host read of the overflow flags when something
        was re-packed (parameter load / update time, never inside a decode loop)."""
        fresh = []
        for n, p in named.items():
            key = (p.data_ptr(), p._version, tuple(p.shape))
            e = self.entries.get(n)
            if e is None or e[0] != key:
                hi, lo, flag = pack_weight(p.detach())
                self.entries[n] = (key, hi, lo, flag)
                fresh.append(n)
        for n in list(self.entries):
            if n not in named:
                del self.entries[n]
        if fresh:
            for n in fresh:
                key, hi, lo, flag = self.entries[n]
                if int(flag.item()) != 0:
                    self.entries[n] = (key, None, None, None)
        akey = tuple((n, e[0]) for n, e in self.entries.items())
        if self.array is None or self.array_key != akey:
            good = [(n, e) for n, e in self.entries.items() if e[1] is not None]
            arr = (_lib.Packed * max(len(good), 1))()
            for i, (n, e) in enumerate(good):
                arr[i] = packed_struct(named[n], e[1], e[2])
            self.array, self.array_key, self.count = arr, akey, len(good)
        return self.array, self.count
