from .losses import PanopticLoss, panoptic_losses

__all__ = ["PanopticLoss", "panoptic_losses"]
