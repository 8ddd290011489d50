from . import partition

__all__ = ["partition"]
