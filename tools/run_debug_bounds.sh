#!/bin/bash
# Builds the library with -DPDC_DEBUG_BOUNDS (device-side asserts on every shared-memory / global index, canary words
# behind every device allocation) into _ab/lib_debug_bounds.so - on the build container, nvcc cross-compiles - or, with
# "run", executes the GPU suite against it on a GPU box: every context is checked for intact canaries when it closes
# (PDC_CHECK_CANARIES=1, srsran_edgeric_5g_b200/capi.py) and an assert that fires fails the launch, hence the test.
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "run" ]; then
  PDC_LIBRARY=$PWD/_ab/lib_debug_bounds.so PDC_CHECK_CANARIES=1 python -m pytest tests -m gpu -x -q \
    --deselect tests/test_gpu_adapters.py 2>&1 | tail -5
  exit 0
fi
mkdir -p _ab
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -cudart static -DPDC_DEBUG_BOUNDS -o _ab/lib_debug_bounds.so srsran_edgeric_5g_b200/csrc/pusch_dec_cuda.cu
echo built _ab/lib_debug_bounds.so
