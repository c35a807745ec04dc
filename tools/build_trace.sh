#!/bin/bash
# Development aid: a second copy of the library with extra compile flags (default: the CGP_UMMA_TRACE time stamps)
#   tools/build_trace.sh [name] [flags...]   ->  tools/_trace/<name>/libconvgp_b200.so
# load it with `tools/f32_check.py --lib tools/_trace/<name>/libconvgp_b200.so [--trace]`.
set -e
cd "$(dirname "$0")/.."
NAME=${1:-trace}; shift || true
FLAGS_EXTRA=${@:--DCGP_UMMA_TRACE}
C=convgp_b200/csrc
O=tools/_trace/$NAME
mkdir -p $O
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --extended-lambda"
nvcc $FLAGS $FLAGS_EXTRA -c $C/api.cu -o $O/api.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $O/libconvgp_b200.so $O/api.o $C/probe.o $C/comm.o $C/linalg.o $C/likelihood.o -lcudart
echo built $O/libconvgp_b200.so
