"""RAM -- robust adaptive Metropolis (Vihola 2012) on the block-per-chain CUDA kernel
(adaptive_mcmc_b200/csrc/ram_block.cuh).  NOT in the reference: BASELINE.json configs[4] asks for
it (synthetic correlated Gaussian d = 200, rank-one Cholesky update dominated); the spec is
SURVEY 8a row 20 and its oracle is oracle/arwmh_numpy.py:ram_step.  Same interface as ARWMH; the
state record is reused: `adapt_state.scale` is the RAM factor S (proposal x' = x + S z),
`loc` / `log_step_size` are carried but unused."""
from __future__ import annotations

from .. import _lib
from .arwmh import ARWMH


class RAM(ARWMH):
    def run_batch(self, batch, num_steps, thinning=1, collect_start=0, collect=("z", "potential_energy"), draws=None,
                  record_accept=False, adapt=True, kernel_kind=_lib.KERNEL_RAM):
        return super().run_batch(batch, num_steps, thinning, collect_start, collect, draws, record_accept, adapt,
                                 kernel_kind)
