from .hashgrid import HashGrid

__all__ = ["HashGrid"]
