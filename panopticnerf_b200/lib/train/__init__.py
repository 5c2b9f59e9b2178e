from .losses import PanopticLoss, panoptic_losses
from .mlp_backward import network_backward, network_forward_autograd, network_forward_rays_autograd, training_step

__all__ = ["PanopticLoss", "panoptic_losses", "network_backward", "network_forward_autograd",
           "network_forward_rays_autograd", "training_step"]
