"""Sampler kernels -- mirror of the reference's ``python/kernels/__init__.py`` for the ARWMH path."""
from .arwmh import ARWMH, ARWMHState, ARWMHAdaptState, ChainBatch, init_to_uniform, init_to_value
from .ram import RAM
from .asss import ASSS, ASSSState, ASSSAdaptState

__all__ = ["ASSS", "ASSSState", "ASSSAdaptState", "RAM", "ARWMH", "ARWMHState", "ARWMHAdaptState", "ChainBatch", "init_to_uniform", "init_to_value"]
