"""Loader-side input compaction: the wire format of the step BEFORE the hot path.

The reference loaders (dataloaders/dataloader_test.py:234-273, dataloaders/dataloader.py:269-303) ship, per image,
  * `obj_dist [37, 1599]` / `pred_dist [65, 21]` fp32 score tables of which the model only takes an arg-max (AttModel.py:374,382-385),
  * `gpn_pool_mtx [.., 37, 37]` diagonal matrices and `att_masks [.., 37]` that only express a sub-graph length,
  * the sub-graph tensors tiled x5 (`seq_per_img`) although the test path reads copy 0 (models/lib/gpn.py:86-94),
  * `gpn_pred_ind` / `gpn_nrel_ind`, which the Sub-GC model never reads (gpn.py:163-170,47).
At 128 images that is 80.5 MB of host->device traffic per step of which 38.8 MB (the region features) are used as such.

`compact_batch` turns the loader tuple into the few tensors the kernels need -- class ids (int16), edge list (uint8), node lists
(uint8) + lengths (uint8), one copy per image -- and `TopDownModel.forward(batch, mode='sample_compact')` decodes from it with
results identical to the loader-shaped call (tests/test_gpu_compact.py).  `needed_only` is the other half of the row: the
reference-signature call with every tensor the kernels do not read replaced by None, so that a caller uploads nothing in vain.
"""
from __future__ import annotations

from dataclasses import dataclass, fields

import torch


@dataclass
class CompactBatch:
    att_feats: torch.Tensor            # f32  [B, N, A]   region features, row N-1 = dummy (zeros)
    obj_cls: torch.Tensor              # i16  [B, N]      1 + argmax_first(obj_dist[:, :, 1:])          (AttModel.py:374)
    pred_cls: torch.Tensor | None      # i16  [B, K]      class of pred_dist as AttModel.py:382-385 takes it (None: not needed)
    rel_ind: torch.Tensor              # u8   [B, K, 2]   (subject, object) node of every edge, pads = N-1
    sub_nodes: torch.Tensor            # u8   [B, 2, M, N] node ids of every sampled sub-graph (both halves), pads = N-1
    sub_len: torch.Tensor              # u8   [B, 2, M]   number of valid nodes (= ones in att_masks = trace of gpn_pool_mtx)

    def to(self, device, non_blocking=False):
        return CompactBatch(*[(getattr(self, f.name).to(device, non_blocking=non_blocking) if getattr(self, f.name) is not None else None)
                              for f in fields(self)])

    def pin_memory(self):
        return CompactBatch(*[(getattr(self, f.name).pin_memory() if getattr(self, f.name) is not None else None) for f in fields(self)])

    def empty_like(self, device):
        return CompactBatch(*[(torch.empty_like(getattr(self, f.name), device=device) if getattr(self, f.name) is not None else None)
                              for f in fields(self)])

    def packed(self, device=None, pin=False, copy=True):
        """The same batch with every field a view into ONE contiguous byte buffer (256-byte aligned fields): a host->device transfer of
        the batch is then a single copy (`copy_` between two packed batches of the same shapes moves the buffer, not the fields)."""
        present = [(f.name, getattr(self, f.name)) for f in fields(self) if getattr(self, f.name) is not None]
        off, spec = 0, []
        for name, t in present:
            nb = t.numel() * t.element_size()
            spec.append((name, t.dtype, tuple(t.shape), off, nb))
            off += (nb + 255) // 256 * 256
        dev = device if device is not None else present[0][1].device
        buf = torch.empty(off, dtype=torch.uint8, device=dev, pin_memory=bool(pin) and torch.device(dev).type == "cpu")
        views = {name: buf[o:o + nb].view(dt).view(shape) for name, dt, shape, o, nb in spec}
        if copy:
            for name, t in present:
                views[name].copy_(t)
        out = CompactBatch(**{f.name: views.get(f.name) for f in fields(self)})
        out._buf, out._layout = buf, tuple((n, str(dt), sh, o) for n, dt, sh, o, _ in spec)
        return out

    def copy_(self, other, non_blocking=False):
        if getattr(self, "_buf", None) is not None and getattr(other, "_buf", None) is not None and self._layout == other._layout:
            self._buf.copy_(other._buf, non_blocking=non_blocking)   # one transfer for the whole batch
            return self
        for f in fields(self):
            a, b = getattr(self, f.name), getattr(other, f.name)
            if a is not None:
                a.copy_(b, non_blocking=non_blocking)
        return self

    def nbytes(self):
        return sum(getattr(self, f.name).numel() * getattr(self, f.name).element_size() for f in fields(self) if getattr(self, f.name) is not None)


def compact_batch(fc_feats, att_feats, att_masks, trip_pred=None, obj_dist=None, obj_box=None, rel_ind=None, pred_fmap=None, pred_dist=None,
                  gpn_obj_ind=None, gpn_pred_ind=None, gpn_nrel_ind=None, gpn_pool_mtx=None, seq_per_img=5, pred_emb_type=1, with_pred=False):
    """Loader tuple (the 13 tensors of AttModel._sample, reference models/AttModel.py:236-237, on the HOST) -> CompactBatch.
    Runs where the loader runs (CPU workers); the arg-max uses the same torch op as the reference."""
    B = att_feats.shape[0]
    obj_cls = (torch.max(obj_dist[:, :, 1:], dim=2)[1] + 1).to(torch.int16)
    pred_cls = None
    if with_pred:
        if pred_emb_type == 1:
            pred_cls = (torch.max(pred_dist[:, :, 1:], dim=2)[1] + 1).to(torch.int16)
        else:
            pred_cls = torch.max(pred_dist, dim=2)[1].to(torch.int16)
    nodes = gpn_obj_ind[::seq_per_img] if gpn_obj_ind.shape[0] == B * seq_per_img else gpn_obj_ind      # copy 0 of every image
    masks = att_masks[::seq_per_img] if att_masks.shape[0] == B * seq_per_img else att_masks
    return CompactBatch(att_feats=att_feats.contiguous().float(), obj_cls=obj_cls.contiguous(), pred_cls=pred_cls,
                        rel_ind=rel_ind.to(torch.uint8).contiguous(), sub_nodes=nodes.to(torch.uint8).contiguous(),
                        sub_len=masks.sum(-1).round().to(torch.uint8).contiguous())


def needed_only(fc_feats, att_feats, att_masks, trip_pred=None, obj_dist=None, obj_box=None, rel_ind=None, pred_fmap=None, pred_dist=None,
                gpn_obj_ind=None, gpn_pred_ind=None, gpn_nrel_ind=None, gpn_pool_mtx=None):
    """The reference-signature argument tuple with None in place of everything the Sub-GC kernels never read (fc_feats is
    overwritten by the sGPN read-out, pred_dist feeds a dead GCN sub-path, gpn_pred_ind / gpn_nrel_ind / gpn_pool_mtx are unused)."""
    return (None, att_feats, att_masks, None, obj_dist, None, rel_ind, None, None, gpn_obj_ind, None, None, None)
