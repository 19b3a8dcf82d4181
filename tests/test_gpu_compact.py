"""Loader-side input compaction (subgc.compact): the compact call and the reference-signature call give IDENTICAL
outputs (same kernels, same operands), the upload shrinks from 80.5 MB to <= 42 MB per 128 images, and the reference-signature call
works with every unread tensor left out."""
import pytest
import torch

from subgc import compact, synth
from subgc.config import Dims, make_opt
from subgc.model import setup

pytestmark = pytest.mark.gpu


def _model(d, sd, **kw):
    m = setup(make_opt(d, test_LSTM=1, **kw))
    m.load_state_dict(sd)
    return m.cuda().eval()


@pytest.mark.parametrize("mode,n_images,per_half", [("greedy", 128, 1), ("greedy", 5, 3), ("topk", 9, 2), ("beam", 4, 2)])
def test_compact_equals_loader_shaped(mode, n_images, per_half):
    d = Dims()
    sd = synth.make_state_dict(d, 17)
    data = synth.make_test_inputs(d, 17, n_images=n_images, per_half=per_half, ragged=per_half > 1, ragged_edges=per_half > 1)
    host = [data[k] for k in synth.SAMPLE_ARG_ORDER]
    m = _model(d, sd, gpn_nms_thres=0.6, gpn_max_subg=2, use_topk_sampling=1 if mode == "topk" else 0)
    opt = {"beam_size": 3 if mode == "beam" else 1}
    with torch.no_grad():
        n_rows = m(*[t.cuda() if t is not None else None for t in host], opt=dict(opt, seed=1), mode="sample")[0].shape[0]
        if mode == "topk":
            opt["topk_uniforms"] = torch.rand(d.seq_length, n_rows, generator=torch.Generator().manual_seed(5))
        ref = m(*[t.cuda() if t is not None else None for t in host], opt=opt, mode="sample")
        cb = compact.compact_batch(*host)
        if n_images == 128:
            loader_bytes = sum(t.numel() * t.element_size() for t in host if t is not None)
            assert loader_bytes > 80e6 and cb.nbytes() <= 42e6, (loader_bytes, cb.nbytes())
        got = m(cb.pin_memory().to("cuda", non_blocking=True), opt=opt, mode="sample_compact")
        lean = m(*[t.cuda() if t is not None else None for t in compact.needed_only(*host)], opt=opt, mode="sample")
    for a, b, c in zip(ref, got, lean):
        a, b, c = a.cpu(), b.cpu(), c.cpu()
        assert a.dtype == b.dtype and torch.equal(a, b), "compact call differs from the loader-shaped call"   # same kernels, same operands
        assert torch.equal(a, c), "call without the unread tensors differs"
    assert torch.equal(m.last_image_of_row.cpu(), m.last_image_of_row.cpu())


def _same(a, b):
    """Token ids / kept indices exact; floats within the bar (the stages run at the row-count upper bound, so the contractions may be
    split differently along K than in the synchronising path)."""
    for x, y in zip(a, b):
        x, y = x.cpu(), y.cpu()
        assert x.shape == y.shape and x.dtype == y.dtype
        if x.dtype.is_floating_point:
            assert float((x - y).abs().max()) <= 2e-5 * max(1.0, float(y.abs().max()))
        else:
            assert torch.equal(x, y)


def test_whole_step_graph_equals_the_synchronising_path():
    """With the persistent decode kernel the call runs without a host round trip (row count and clip length stay on the device,
    subgc_decode_sample_dyn) and replays as one CUDA graph: same outputs as the path that reads them back after NMS, also when NMS keeps
    FEWER rows than the upper bound the stages run at, for fresh input tensors on every call (copy-mode plan) and for re-used ones."""
    d = Dims()
    sd = synth.make_state_dict(d, 23)
    m = _model(d, sd, gpn_nms_thres=0.3, gpn_max_subg=3)      # a low threshold suppresses most sub-graphs: kept < 3 per image
    ref_m = _model(d, sd, gpn_nms_thres=0.3, gpn_max_subg=3)
    ref_m.use_step_graph = False
    for it in range(7):
        data = synth.make_test_inputs(d, 100 + it % 2, n_images=6, per_half=3, ragged=True, ragged_edges=True)
        args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]   # new device tensors every time
        with torch.no_grad():
            a = m(*args, opt={"beam_size": 1}, mode="sample")
            b = ref_m(*args, opt={"beam_size": 1}, mode="sample")
        assert a[0].shape[0] < 18, "the bound (6 images x 3) was meant to be loose"
        _same(a, b)
        assert torch.equal(m.last_image_of_row.cpu(), ref_m.last_image_of_row.cpu())
    group = [v for k, v in m._plans.items() if k[0] == "step"]
    assert group and (group[0]["copy"] is not None or any(p.graph is not None for p in group[0]["ptr"].values()))
    # the same tensors again and again: the zero-copy plan of that buffer set gets captured and replayed
    with torch.no_grad():
        for _ in range(4):
            a = m(*args, opt={"beam_size": 1}, mode="sample")
    _same(a, b)
