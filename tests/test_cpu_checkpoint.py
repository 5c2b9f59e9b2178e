"""CPU: reference-layout checkpoints ({'net': state_dict with wrapper prefixes, 'epoch': ...}) load into
the product Network."""
import torch

import panopticnerf_b200 as PN
from oracle import reference_renderer as O
from panopticnerf_b200 import synthetic as S
from panopticnerf_b200.lib.utils import net_utils


def test_load_reference_layout(tmp_path):
    cfg = PN.make_cfg("cfg3")
    ref = S.init_network_weights(O.make_network(cfg), seed=3)
    wrapped = {"net." + k: v for k, v in ref.state_dict().items()}
    torch.save({"net": wrapped, "optim": {}, "scheduler": {}, "recorder": {}, "epoch": 17}, tmp_path / "17.pth")
    torch.save({"net": {"module.net." + k: v for k, v in ref.state_dict().items()}, "epoch": 20}, tmp_path / "latest.pth")
    net = PN.make_network(cfg)
    assert net_utils.load_network(net, str(tmp_path), epoch=17) == 17
    for k, v in ref.state_dict().items():
        assert torch.equal(net.state_dict()[k], v), k
    net2 = PN.make_network(cfg)
    assert net_utils.load_network(net2, str(tmp_path)) == 20          # latest.pth wins for epoch=-1
    assert torch.equal(net2.rgb_linear.weight, ref.rgb_linear.weight)


def test_missing_and_partial(tmp_path):
    cfg = PN.make_cfg("cfg3")
    small = O.make_network(PN.make_cfg("cfg2"))                          # no heads
    torch.save({"net": small.state_dict()}, tmp_path / "0.pth")
    net = PN.make_network(cfg)
    try:
        net_utils.load_network(net, str(tmp_path), epoch=0)
        assert False, "strict load of a head-less checkpoint must fail"
    except KeyError as e:
        assert "semantic_linears" in str(e) or "instance_linears" in str(e)
    net_utils.load_network(net, str(tmp_path), epoch=0, strict=False)    # trunk loads, heads keep their init
    assert torch.equal(net.pts_linears[3].weight, small.pts_linears[3].weight)
    try:
        net_utils.load_network(net, str(tmp_path / "nope"))
        assert False
    except FileNotFoundError:
        pass
