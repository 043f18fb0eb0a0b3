"""Drop-in alias: the reference exports ``PLDA`` and ``LDA`` from a package called
``liblda`` (``python/liblda/__init__.py:1-3``); existing user code keeps working."""
from plda_b200 import LDA, PLDA

__all__ = ["PLDA", "LDA"]
