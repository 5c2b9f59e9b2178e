"""CPU tier: the closed form of the compositing backward pass (oracle.raw2outputs_backward - the algorithm the
CUDA kernel implements) against torch.autograd through the oracle's forward raw2outputs."""
import pytest
import torch

from oracle import reference_renderer as O
from util import assert_close, rms


def _case(R, N, C, K, seed, boxes):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(R, N, 4 + C + K, generator=g)
    raw[..., 3] = raw[..., 3] * 0.6 - 0.1
    z = torch.sort(torch.rand(R, N, generator=g) * 58 + 2, -1).values
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * (0.5 + torch.rand(R, 1, generator=g))
    kw = {}
    if boxes:
        kw = {"sample_box": torch.randint(-1, 9, (R, N), generator=g, dtype=torch.int32),
              "box_sem": torch.randint(0, max(C, 1), (9,), generator=g, dtype=torch.int32),
              "box_inst": torch.randint(0, max(K, 1), (9,), generator=g, dtype=torch.int32)}
    shapes = {"rgb_map": (R, 3), "depth_map": (R,), "acc_map": (R,), "weights": (R, N), "semantic_map": (R, C),
              "instance_map": (R, K), "fixed_semantic_map": (R, C), "fixed_instance_map": (R, K)}
    keys = ["rgb_map", "depth_map", "acc_map", "weights"] + (["semantic_map"] if C else []) + (["instance_map"] if K else [])
    if boxes:
        keys += (["fixed_semantic_map"] if C else []) + (["fixed_instance_map"] if K else [])
    ups = {k: torch.randn(*shapes[k], generator=g) for k in keys}
    return raw, z, d, kw, ups


@pytest.mark.parametrize("R,N,C,K,boxes,white,mask", [
    (40, 64, 0, 0, False, False, False), (33, 33, 7, 5, True, True, True), (8, 1, 3, 0, False, True, False),
    (16, 192, 45, 50, True, False, True)])
def test_closed_form_matches_autograd(R, N, C, K, boxes, white, mask):
    raw, z, d, kw, ups = _case(R, N, C, K, seed=R * N + C, boxes=boxes)
    x = raw.clone().requires_grad_(True)
    out = O.raw2outputs(x, z, d, num_classes=C, num_instances=K, white_bkgd=white, mask_outside=mask, **kw)
    (ref,) = torch.autograd.grad(sum((out[k] * u).sum() for k, u in ups.items()), x)
    got = O.raw2outputs_backward(raw, z, d, ups, white_bkgd=white, num_classes=C, num_instances=K,
                                 mask_outside=mask, **kw)
    assert torch.isfinite(ref).all()
    for name, sl in (("rgb", slice(0, 3)), ("sigma", slice(3, 4)), ("sem", slice(4, 4 + C)), ("inst", slice(4 + C, 4 + C + K))):
        if ref[..., sl].numel():
            assert_close(got[..., sl], ref[..., sl], max(rms(ref[..., sl]), 1e-12), name)


def test_empty_and_unoccupied_samples_get_no_density_gradient():
    raw, z, d, kw, ups = _case(10, 16, 0, 0, seed=1, boxes=False)
    raw[..., 3] = -raw[..., 3].abs()                      # sigma_raw <= 0 everywhere: relu is flat
    got = O.raw2outputs_backward(raw, z, d, ups)
    assert bool((got[..., 3] == 0).all())
    assert bool((got[..., :3] == 0).all())                # and all weights are zero, so no colour gradient either
