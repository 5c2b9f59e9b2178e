from .panoptic import fuse_panoptic

__all__ = ["fuse_panoptic"]
