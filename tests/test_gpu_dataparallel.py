"""nn.DataParallel over LossWrapper, as the reference's training driver wraps it (train.py:96-98: `dp_lw_model = DataParallel(lw_model)`,
train.py:151-162: losses averaged over the replicas, one backward).  Replicas of one module run in one Python thread per device and are
shallow copies of its __dict__: everything the model caches (scratch, weight tables, plans) is therefore keyed by device
(subgc.model._DeviceState) and a replica's weights are fetched by attribute path (it has no registered Parameters).  Needs 2 GPUs."""
import pytest
import torch

from helpers import rel_err, t2n
from subgc import synth
from subgc.config import SMALL, make_opt
from subgc.model import LossWrapper, setup

pytestmark = pytest.mark.gpu


def _call(data):
    return (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
            data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_loss_wrapper_under_dataparallel_matches_per_shard_mean():
    d = SMALL
    sd = synth.make_state_dict(d, 9, logit_gain=4.0)
    host = synth.make_train_inputs(d, 9, n_images=4, gpn_batch=2)
    data = {k: (v.cuda(0) if torch.is_tensor(v) else v) for k, v in host.items()}
    model = setup(make_opt(d))
    model.load_state_dict(sd)
    model.cuda(0).train()
    model.dropout_enabled = False
    lw = LossWrapper(model, None)
    dp = torch.nn.DataParallel(lw, device_ids=[0, 1])
    for it in range(2):   # twice: the replicas of the second call find the per-device state of the first
        for p in model.parameters():
            p.grad = None
        out = dp(*_call(data))
        assert out["lang_loss"].shape == (2,) and out["gpn_loss"].shape == (2,)
        loss = out["lang_loss"].mean() + out["gpn_loss"].mean()          # train.py:154-158
        loss.backward()
    got = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert set(model._states) == {0, 1}, "every device must have its own cached state"
    # the same two shards, one after the other, on one device
    shard_grads, shard_losses = [], []
    B = 4
    for s in range(2):
        sl = {k: (v[(s * B // 2) * (v.shape[0] // B):((s + 1) * B // 2) * (v.shape[0] // B)] if torch.is_tensor(v) else v) for k, v in data.items()}
        for p in model.parameters():
            p.grad = None
        o = lw(*_call(sl))
        (o["lang_loss"] + o["gpn_loss"]).backward()
        shard_losses.append((float(o["lang_loss"]), float(o["gpn_loss"])))
        shard_grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    for s in range(2):
        assert abs(float(out["lang_loss"][s]) - shard_losses[s][0]) <= 1e-5 * max(1.0, abs(shard_losses[s][0]))
        assert abs(float(out["gpn_loss"][s]) - shard_losses[s][1]) <= 1e-5
    assert set(shard_grads[0]) <= set(got)
    for n in got:
        if n not in shard_grads[0]:   # parameters of the dead GCN sub-path: no gradient on one device, zeros through DataParallel's broadcast
            assert float(got[n].abs().sum()) == 0.0, n
            continue
        ref = 0.5 * (shard_grads[0][n] + shard_grads[1][n])
        assert rel_err(t2n(got[n]), t2n(ref)) <= 2e-5 or float(ref.abs().max()) < 1e-12, n


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_inference_on_two_devices_from_one_process():
    """One model object per device in one process (the per-device state must not leak between them): both decode the same captions."""
    from subgc.config import Dims
    d = Dims()
    sd = synth.make_state_dict(d, 3)
    data = synth.make_test_inputs(d, 3, n_images=4, per_half=2, ragged=True, ragged_edges=True)
    outs = []
    for dev in (0, 1):
        m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.6, gpn_max_subg=2))
        m.load_state_dict(sd)
        m.cuda(dev).eval()
        args = [data[k].cuda(dev) if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
        with torch.cuda.device(dev), torch.no_grad():
            for _ in range(3):
                r = m(*args, opt={"beam_size": 1}, mode="sample")
        outs.append([t.cpu() for t in r])
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][3], outs[1][3])
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 2e-5 * 10
