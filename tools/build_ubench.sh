#!/bin/bash
# builds the sm_100a micro-benchmarks (tools/_bin/ is git-ignored; the binary travels with gpurun)
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  tools/ubench_sm100.cu refign_b200/csrc/tensormap.cu -o tools/_bin/ubench_sm100
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  tools/trace_attn_bwd.cu refign_b200/csrc/tensormap.cu -o tools/_bin/trace_attn_bwd
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  tools/trace_attn_fwd.cu refign_b200/csrc/tensormap.cu -o tools/_bin/trace_attn_fwd
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  tools/trace_gemm.cu refign_b200/csrc/tensormap.cu -o tools/_bin/trace_gemm
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  tools/trace_local_corr.cu refign_b200/csrc/tensormap.cu -o tools/_bin/trace_local_corr
