from .gaussian_moment_descriptor import GaussianMomentDescriptor

__all__ = ["GaussianMomentDescriptor"]
