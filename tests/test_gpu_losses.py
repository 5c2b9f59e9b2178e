"""GPU parity of the loss kernel (loss side): `pnr_losses` through the autograd node vs the
oracle's torch losses - the four terms and, via autograd on both sides, the gradients w.r.t. every map."""
import pytest
import torch

from oracle import reference_losses as OL
from panopticnerf_b200.lib.train import PanopticLoss, panoptic_losses
from util import assert_close, rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(R, C, seed, prob):
    g = torch.Generator().manual_seed(seed)
    rgb, rgb0, gt = torch.rand(R, 3, generator=g), torch.rand(R, 3, generator=g), torch.rand(R, 3, generator=g)
    depth = torch.rand(R, generator=g) * 50 + 1
    depth_gt = torch.where(torch.rand(R, generator=g) > 0.3, depth + torch.randn(R, generator=g), torch.zeros(R))
    sem = torch.rand(R, C, generator=g) if prob else torch.randn(R, C, generator=g) * 3
    if prob:
        sem = sem / sem.sum(1, keepdim=True) * torch.rand(R, 1, generator=g)
    fix = torch.rand(R, C, generator=g) * (torch.rand(R, C, generator=g) > 0.7)          # many exact zeros -> the eps branch
    label = torch.randint(-1, C, (R,), generator=g)
    conf = torch.rand(R, generator=g)
    return rgb, rgb0, depth, sem, fix, gt, depth_gt, label, conf


@pytest.mark.parametrize("R,C,prob", [(1000, 45, False), (257, 19, True), (33, 3, False), (4096, 100, True)])
def test_losses_and_gradients_match_autograd(R, C, prob):
    rgb, rgb0, depth, sem, fix, gt, depth_gt, label, conf = _case(R, C, R + C, prob)
    w = (1.0, 0.1, 0.5, 2.0)
    maps_ref = [t.clone().requires_grad_(True) for t in (rgb, rgb0, depth, sem, fix)]
    tot_ref, terms_ref = OL.losses(*maps_ref, gt, depth_gt, label, conf, w, prob, 1e-6)
    tot_ref.backward()
    maps = [t.clone().to(DEV).requires_grad_(True) for t in (rgb, rgb0, depth, sem, fix)]
    tot, terms = PanopticLoss.apply(*maps, gt.to(DEV), depth_gt.to(DEV), label.to(DEV), conf.to(DEV), w, prob, 1e-6)
    (tot * 3.0).backward()                                   # an upstream factor must scale the gradients
    assert torch.allclose(terms.cpu(), terms_ref.detach(), rtol=2e-5, atol=1e-6)
    assert float(tot) == pytest.approx(float(tot_ref), rel=2e-5)
    assert not terms.requires_grad
    for m, mr, name in zip(maps, maps_ref, ("rgb", "rgb0", "depth", "sem", "fix")):
        assert_close(m.grad.cpu() / 3.0, mr.grad, rms(mr.grad), f"d/d{name}", rel=1e-4)


def test_losses_subset_of_terms_and_ignored_rays():
    rgb, rgb0, depth, sem, fix, gt, depth_gt, label, conf = _case(300, 7, 5, False)
    out = {"rgb_map": rgb.to(DEV), "semantic_map": sem.to(DEV)}
    total, terms = panoptic_losses(out, {"rgb": gt.to(DEV), "pseudo_label": torch.full((300,), -1)}, (1.0, 1.0, 1.0, 1.0))
    tot_ref, terms_ref = OL.losses(rgb, None, None, sem, None, gt, None, torch.full((300,), -1))
    assert float(terms["sem"]) == 0.0 and float(terms["depth"]) == 0.0 and float(terms["fix"]) == 0.0
    assert float(total) == pytest.approx(float(tot_ref), rel=2e-5) and float(terms["rgb"]) == pytest.approx(float(terms_ref[0]), rel=2e-5)


def test_loss_backpropagates_into_raw_through_the_compositing_node():
    """loss -> maps (pnr_losses) -> raw (pnr_composite_backward): dL/draw equals autograd through the oracle's
    raw2outputs + losses."""
    from oracle import reference_renderer as O
    from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P
    g = torch.Generator().manual_seed(2)
    R, N, C = 200, 64, 9
    raw = torch.randn(R, N, 4 + C, generator=g)
    raw[..., 3] = raw[..., 3] * 0.6 - 0.1
    z = torch.sort(torch.rand(R, N, generator=g) * 40 + 2, -1).values
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    gt, depth_gt = torch.rand(R, 3, generator=g), torch.rand(R, generator=g) * 40
    label = torch.randint(-1, C, (R,), generator=g)
    w = (1.0, 0.05, 0.3, 0.0)
    rr = raw.clone().requires_grad_(True)
    o = O.raw2outputs(rr, z, d, num_classes=C)
    OL.losses(o["rgb_map"], None, o["depth_map"], o["semantic_map"], None, gt, depth_gt, label, None, w)[0].backward()
    rg = raw.clone().to(DEV).requires_grad_(True)
    og = P.raw2outputs_autograd(rg, z.to(DEV), d.to(DEV), num_classes=C)
    total, _ = PanopticLoss.apply(og["rgb_map"], None, og["depth_map"], og["semantic_map"], None, gt.to(DEV),
                                  depth_gt.to(DEV), label.to(DEV), None, w, False, 1e-8)
    total.backward()
    for name, sl in (("rgb", slice(0, 3)), ("sigma", slice(3, 4)), ("sem", slice(4, 4 + C))):
        assert_close(rg.grad[..., sl].cpu(), rr.grad[..., sl], rms(rr.grad[..., sl]), f"dL/draw[{name}]", rel=1e-4)
