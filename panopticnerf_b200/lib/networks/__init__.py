"""The reference's ``lib.networks`` plugin surface: make_network / make_renderer."""
from .make_network import make_network
from .renderer.make_renderer import make_renderer

__all__ = ["make_network", "make_renderer"]
