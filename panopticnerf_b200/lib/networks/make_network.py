"""make_network(cfg) -> Network   (reference: lib/networks/make_network.py, not in
the mount).  The reference loads ``cfg.network_module`` with the removed ``imp`` module; here the same
string keys are resolved with importlib."""
from __future__ import annotations

import importlib


def make_network(cfg):
    module = getattr(cfg, "network_module", "panopticnerf_b200.lib.networks.panopticnerf.network")
    return importlib.import_module(module).Network(cfg)
