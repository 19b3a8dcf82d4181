"""Fused global-norm clip + Adam against the reference's sequence: utils.clip_gradient_norm (misc/utils.py:174-200,
restated below line by line) + torch.optim.Adam.step() (misc/utils.py:236), three steps, parameters / moments within 1e-6 relative."""
import pytest
import torch

from subgc.optim import ClipAdam

pytestmark = pytest.mark.gpu


def clip_gradient_norm(optimizer, clip_norm=10.):
    totalnorm = 0
    for group in optimizer.param_groups:
        for p in group['params']:
            if p.requires_grad and p.grad is not None:
                modulenorm = p.grad.data.norm(2)
                totalnorm += modulenorm ** 2
    totalnorm = totalnorm ** (1. / 2)
    norm = clip_norm / max(totalnorm, clip_norm)
    for group in optimizer.param_groups:
        for p in group['params']:
            if p.requires_grad and p.grad is not None:
                p.grad.mul_(norm)
    return totalnorm


@pytest.mark.parametrize("grad_scale,wd", [(0.01, 0.0), (30.0, 0.0), (3.0, 1e-3)])
def test_clip_adam_matches_reference_sequence(grad_scale, wd):
    g = torch.Generator().manual_seed(0)
    shapes = [(9488, 1000), (4000, 3000), (1000,), (1,), (7, 13), (512, 2048), (3,), (16385,), (4000, 1000), (37, 5, 3)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    frozen = torch.nn.Parameter(torch.randn(5, 5).cuda())            # never gets a gradient (the dead GCN units of Sub-GC)
    ref = torch.optim.Adam(ref_p + [frozen], 5e-4, (0.9, 0.999), 1e-8, weight_decay=wd)
    ours = ClipAdam(our_p + [frozen], 5e-4, (0.9, 0.999), 1e-8, weight_decay=wd, clip_norm=10.0, write_grad=True)
    for it in range(3):
        if it == 1:
            for grp in ref.param_groups + ours.param_groups:            # utils.set_lr between steps (train.py:112-124)
                grp["lr"] = 2.5e-4
        for a, b in zip(ref_p, our_p):
            gr = torch.randn(a.shape, generator=g).cuda() * grad_scale
            a.grad = gr.clone()
            b.grad = gr.clone()
        tn = clip_gradient_norm(ref, 10.)
        ref.step()
        ours.step()
        assert abs(float(ours.norm[0]) - float(tn)) <= 1e-5 * float(tn)
        for a, b in zip(ref_p, our_p):
            sa, sb = ref.state[a], ours.state[b]
            for x, y in ((a, b), (sa["exp_avg"], sb["exp_avg"]), (sa["exp_avg_sq"], sb["exp_avg_sq"]), (a.grad, b.grad)):
                diff = (x.detach() - y.detach()).abs()
                tol = 1e-6 * max(1e-3, float(x.detach().abs().max()))
                if wd == 0.0:
                    assert float(diff.max()) <= tol, (it, tuple(a.shape), float(diff.max()))
                else:
                    # with weight decay g + wd * p cancels for a few elements; there Adam's update lr * g / (|g| + eps) amplifies the
                    # 1e-7 relative difference of the two clip coefficients (fp32 vs fp64 norm accumulation): bounded by the step size
                    assert float((diff > tol).float().mean()) <= 1e-5 and float(diff.max()) <= 5e-4, (it, tuple(a.shape), float(diff.max()))
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and sd["param_groups"][0]["lr"] == 2.5e-4
    assert frozen.grad is None and not ours.state[frozen]
