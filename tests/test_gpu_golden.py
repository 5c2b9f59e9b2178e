"""GPU vs the committed golden fixtures (no oracle import at run time): every stage through the C ABI,
the fused MLP, and a config-1 render."""
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

import panopticnerf_b200 as PN
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P
from util import assert_close, check_render_outputs, rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = torch.load(Path(__file__).parent / "golden" / "render_golden.pt", weights_only=False)
d = lambda t: t.to(DEV)


def test_stage_fixtures_bit_exact():
    s = G["stage"]
    bx = s["boxes"]
    hit, bid, tin, tout = P.intersect(d(s["rays"]), d(bx["box_center"]), d(bx["box_half"]), d(bx["box_rot"]), 4)
    assert torch.equal(hit.cpu(), s["hit"]) and torch.equal(bid.cpu(), s["box_id"])
    assert torch.equal(tin.cpu(), s["t_in"]) and torch.equal(tout.cpu(), s["t_out"])
    from panopticnerf_b200 import synthetic as S
    near, far = P.scene_near_far(d(s["rays"]), torch.tensor(S.SCENE_AABB), 0.05, 80.0)
    assert torch.equal(near.cpu(), s["near"]) and torch.equal(far.cpu(), s["far"])
    z, sb = P.stratified_z(near, far, d(s["t_vals"]), 1.0, d(s["u"]), bid, tin, tout, want_tags=True)
    assert torch.equal(z.cpu(), s["z"]) and torch.equal(sb.cpu(), s["sample_box"])
    z_f, z_all, idx = P.sample_pdf(z, d(s["weights"]), 32, u=d(s["u_fine"]), want_idx=True)
    assert torch.equal(idx.cpu(), s["idx"]) and torch.equal(z_f.cpu(), s["z_fine"]) and torch.equal(z_all.cpu(), s["z_all"])


def test_composite_and_embed_fixtures():
    c, s = G["composite"], G["stage"]
    out = P.raw2outputs(d(c["raw"]), d(s["z"]), d(s["rays"][:, 3:].contiguous()), num_classes=5, num_instances=6,
                        sample_box=d(s["sample_box"]), box_sem=d(c["box_sem"]), box_inst=d(c["box_inst"]))
    floors = {"rgb_map": 1e-2, "acc_map": 1e-3, "weights": 1e-3, "depth_map": 0.4, "disp_map": 1.0 / 40}
    for k, v in c["out"].items():
        assert_close(out[k], v, floors.get(k, max(rms(v), 1e-6)), k)
    e = G["embed"]
    assert_close(P.embed(d(e["x"]), 10), e["e10"], 1.0, "embed10", rel=2e-6)
    assert_close(P.embed(d(e["x"]), 4), e["e4"], 1.0, "embed4", rel=2e-6)


def test_network_and_render_fixture():
    n = G["net"]
    cfg = SimpleNamespace(**n["cfg"])
    net = PN.make_network(cfg)
    net.load_state_dict(n["state"])
    net = net.to(DEV)
    with torch.no_grad():
        raw = net(d(n["pts"]), d(n["viewdirs"]))
    for name, sl in (("rgb", slice(0, 3)), ("sigma", slice(3, 4)), ("sem", slice(4, 9)), ("inst", slice(9, 15))):
        assert_close(raw[:, sl], n["raw"][:, sl], rms(n["raw"][:, sl]), name)
    out = PN.make_renderer(cfg, net).render({k: d(v) for k, v in n["batch"].items()})
    for k in ("hit_mask", "box_id", "z_vals_0"):
        assert torch.equal(out[k].cpu().to(n["render"][k].dtype), n["render"][k]), k
    check_render_outputs(out, {k: v for k, v in n["render"].items() if k not in ("hit_mask", "box_id", "z_vals_0")},
                         float(n["render"]["far"].max()))
