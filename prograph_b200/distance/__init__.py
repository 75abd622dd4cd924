"""Pluggable distance functions: ``fn(X, Y, similarity=False) -> (M, N)``
(the protocol of prograph/distance/__init__.py:1-3 and README.md:48)."""
from .hamming import hamming
from .minkowski import minkowski
from .utils import clean_input

__all__ = ["hamming", "minkowski", "clean_input"]
