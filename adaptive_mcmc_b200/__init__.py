"""adaptive_mcmc_b200 -- B200-native many-chain adaptive Metropolis sampler.

Drop-in for the sampling loop of savelovme/adaptive-mcmc (``python/kernels`` ARWMH + the
drivers that call it), built from hand-written sm_100a CUDA kernels behind a C ABI
(``include/amcmc.h`` -> ``libamcmc.so``).  No CPU fallback, no Triton, no multi-backend dispatch.
"""
from . import models
from .kernels import ARWMH, RAM, ASSS, ASSSState, ASSSAdaptState, ARWMHState, ARWMHAdaptState, ChainBatch, init_to_uniform, init_to_value
from .infer import MCMC
from .utils.kernel_utils import ns_logscale, concat_trees, collect_states_logscale
from . import diagnostics
from .custom import custom_model

__all__ = [
    "models", "ARWMH", "RAM", "ASSS", "ASSSState", "ASSSAdaptState", "ARWMHState", "ARWMHAdaptState", "ChainBatch", "init_to_uniform", "init_to_value",
    "custom_model", "MCMC", "ns_logscale", "concat_trees", "collect_states_logscale", "diagnostics",
]
