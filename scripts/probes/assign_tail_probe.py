"""Timing builds of assign.cu only (-DAMCMC_ASSIGN_TIMING [-DAMCMC_ASSIGN_FIXROW]): the solver may not terminate correctly
with FIXROW (rows aliased onto 64 L2-resident ones), so bound the work and only read the [tail] lines."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from adaptive_mcmc_b200.utils import evaluation as ev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, 10, device="cuda", generator=g)
y = torch.randn(n, 10, device="cuda", generator=g) * 1.05 + 0.02
cm = ev.cost_matrix(x, y, 2.0)
try:
    ev.linear_sum_assignment(cm)
except Exception as e:  # noqa: BLE001
    print("solver:", e)
torch.cuda.synchronize()
