"""gpclust_b200 -- B200-native (sm_100a) implementation of GPclust's collapsed variational-Bayes hot path.

Drop-in for `from GPclust import MOG, MOHGP, OMGP` on that path (reference: GPclust/__init__.py:5-7).
Importing the package loads the C-ABI library (gpclust_b200/_lib/libgpclust_b200.so); a missing library
raises -- there is no CPU fallback.
"""
from . import _cabi  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import kern  # noqa: F401
from .MOG import MOG  # noqa: F401
from .MOHGP import MOHGP  # noqa: F401
from .OMGP import OMGP  # noqa: F401

__all__ = ["MOG", "MOHGP", "OMGP", "kern"]
