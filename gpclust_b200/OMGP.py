"""Host mirror of GPclust/OMGP.py (class OMGP): overlapping mixtures of Gaussian processes.

Placeholder until the dense per-component path lands.
"""
from .collapsed_mixture import CollapsedMixture


class OMGP(CollapsedMixture):
    def __init__(self, X, Y, K=2, kernels=None, variance=1.0, alpha=1.0, prior_Z="symmetric", name="OMGP"):
        raise NotImplementedError("OMGP: the dense per-component device path is not built yet")
