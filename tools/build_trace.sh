#!/bin/bash
# Development aid: a second copy of the library with the CGP_UMMA_TRACE time stamps compiled in
# (tools/_trace/libconvgp_b200.so; load it with `tools/f32_check.py --lib tools/_trace/libconvgp_b200.so --trace`).
set -e
cd "$(dirname "$0")/.."
C=convgp_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --extended-lambda"
nvcc $FLAGS -DCGP_UMMA_TRACE ${CGP_TRACE_EXTRA:-} -c $C/api.cu -o tools/_trace/api.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/_trace/libconvgp_b200.so tools/_trace/api.o $C/probe.o $C/comm.o $C/linalg.o $C/likelihood.o -lcudart
echo built tools/_trace/libconvgp_b200.so
