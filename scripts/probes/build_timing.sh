#!/bin/bash
# Rebuild libamcmc.so with the phase clocks of the adaptive tensor-core kernel compiled in (-DAMCMC_TC_TIMING).
# Run `touch adaptive_mcmc_b200/csrc/diamonds_tc_adapt.cu && make -C adaptive_mcmc_b200/csrc` afterwards to go back.
set -e
cd "$(dirname "$0")/../../adaptive_mcmc_b200/csrc"
make -j"$(nproc)" >/dev/null 2>&1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -DAMCMC_TC_TIMING $TCA_EXTRA -c diamonds_tc_adapt.cu -o diamonds_tc_adapt.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libamcmc.so *.o -lcudart -ldl
