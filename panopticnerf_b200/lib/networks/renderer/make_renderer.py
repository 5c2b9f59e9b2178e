"""make_renderer(cfg, network) -> Renderer   (reference:
lib/networks/renderer/make_renderer.py, not in the mount)."""
from __future__ import annotations

import importlib


def make_renderer(cfg, network, network_fine=None):
    module = getattr(cfg, "renderer_module",
                     "panopticnerf_b200.lib.networks.renderer.panopticnerf_renderer")
    return importlib.import_module(module).Renderer(cfg, network, network_fine)
