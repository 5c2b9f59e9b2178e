from .losses import PanopticLoss, panoptic_losses
from .mlp_backward import network_backward, network_forward_autograd

__all__ = ["PanopticLoss", "panoptic_losses", "network_backward", "network_forward_autograd"]
