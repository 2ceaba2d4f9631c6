#!/bin/bash
# Builds the library with -DH2_PHASE_TIMING into tools/_build/ (here, before gpurun) or runs tools/phase_probe.py against
# it (on the GPU box): tools/phase_probe.sh build | run
set -e
cd "$(dirname "$0")/.."
case "$1" in
  build)
    mkdir -p tools/_build
    env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
      -cudart static -DH2_PHASE_TIMING -o tools/_build/libpdc_phase.so srsran_edgeric_5g_b200/csrc/pusch_dec_cuda.cu
    ;;
  run)
    PDC_LIBRARY=$PWD/tools/_build/libpdc_phase.so python tools/phase_probe.py
    ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
