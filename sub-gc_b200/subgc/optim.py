"""Fused global-norm clip + Adam: `utils.clip_gradient_norm(optimizer, 10.)` followed by `optimizer.step()` of the
reference's training loop (misc/utils.py:174-200,236; train.py:107-124,163-164) as ONE call that launches three kernels and never
synchronises with the host.

`ClipAdam` is a torch.optim.Optimizer: param_groups / state_dict() have torch.optim.Adam's layout (state[p] = step, exp_avg,
exp_avg_sq), so `utils.set_lr`, `optimizer.state_dict()` / `load_state_dict()` of the reference's checkpoints keep working.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=10.0, write_grad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.clip_norm = float(clip_norm)
        self.write_grad = bool(write_grad)
        self._tables = {}      # per param group: (key, device chunk table, partial scratch)
        self.norm = None       # device [2]: total gradient norm, clip coefficient of the last step

    def _table(self, gi, group, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(), p.numel())
                    for p in plist)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent
        chunk = lib().subgc_opt_chunk_elems()
        rows = []
        for p in plist:
            st = self.state[p]
            n = p.numel()
            for o in range(0, n, chunk):
                rows.append((p.data_ptr() + 4 * o, p.grad.data_ptr() + 4 * o, st["exp_avg"].data_ptr() + 4 * o, st["exp_avg_sq"].data_ptr() + 4 * o,
                             min(chunk, n - o)))
        tab = np.zeros(len(rows), dtype=[("p", "u8"), ("g", "u8"), ("m", "u8"), ("v", "u8"), ("n", "i4"), ("pad", "i4")])
        for i, r in enumerate(rows):
            tab[i] = (r[0], r[1], r[2], r[3], r[4], 0)
        dev = plist[0].device
        table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        ent = (key, table, torch.empty(len(rows), device=dev), len(rows))
        self._tables[gi] = ent
        return ent

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.requires_grad and p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("ClipAdam needs contiguous fp32 CUDA parameters and gradients (there is no CPU path)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {int(self.state[p]["step"]) for p in plist}
            if len(steps) != 1:
                raise RuntimeError("ClipAdam: parameters of one group must share the step count")
            t = steps.pop() + 1
            key, table, partial, n_chunks = self._table(gi, group, plist)
            if self.norm is None or self.norm.device != table.device:
                self.norm = torch.zeros(2, device=table.device)
            b1, b2 = group["betas"]
            check(lib().subgc_clip_adam_step(ptr(table), n_chunks, self.clip_norm, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                             float(group["weight_decay"]), t, int(self.write_grad), ptr(partial), ptr(self.norm),
                                             torch.cuda.current_stream().cuda_stream), "subgc_clip_adam_step")
            for p in plist:
                self.state[p]["step"] = t
        return loss
