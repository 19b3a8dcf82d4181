"""Runs the unmodified reference from baseline/_ref (bench.py --impl reference).  The reference files are untouched; what it cannot
find offline is supplied from outside, exactly the three shims SURVEY §8c lists:
  1. misc.utils.obj_edge_vectors -> seeded N(0,1) vectors (data/glove.6B.300d.pt cannot be downloaded; every weight is overwritten by
     load_state_dict anyway),
  2. obj_name_path / rel_name_path point at the bundled class-name tables,
  3. torch.Tensor.cuda -> identity, because models/CaptionModel.py:129,171 call .cuda() unconditionally inside beam search.
The model is driven as the reference can be driven: ONE image per call (models/lib/gpn.py:84 asserts it), CPU, fp32."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "models", "AttModel.py"))


def load(dims, sd, make_opt, **opt_over):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import misc.utils as ref_utils

    def fake_glove(names, wv_type="glove.6B", wv_dir="data/", wv_dim=300):
        return torch.randn(len(names), wv_dim, generator=torch.Generator().manual_seed(len(names)))

    ref_utils.obj_edge_vectors = fake_glove
    import models as ref_models
    import models.AttModel as ref_att
    ref_att.obj_edge_vectors = fake_glove
    torch.Tensor.cuda = lambda self, *a, **k: self
    opt = make_opt(dims, **opt_over)
    opt.obj_name_path = os.path.join(REF, "data/object_names_1600-0-20.npy")
    opt.rel_name_path = os.path.join(REF, "data/predicate_names_1600-0-20.npy")
    model = ref_models.setup(opt)
    model.load_state_dict(sd, strict=True)
    return model.eval()


def image_args(data, names, b, seq_per_img=5):
    """The loader tuple of image b (dataloaders/dataloader_test.py:191-203): [1, ...] image tensors, [5, ...] sub-graph tensors."""
    out = []
    for k in names:
        t = data[k]
        if t is None:
            out.append(None)
        elif t.shape[0] == data["att_feats"].shape[0]:
            out.append(t[b:b + 1])
        else:
            out.append(t[seq_per_img * b:seq_per_img * (b + 1)])
    return out


def run(model, data, names, images, opt):
    """One pass over `images` (one reference call each); returns (captions, seconds, [seq tensors])."""
    seqs = []
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in images:
            r = model(*image_args(data, names, b), opt=dict(opt), mode="sample")
            seqs.append(r[0])
    return sum(s.shape[0] for s in seqs), time.perf_counter() - t0, seqs
