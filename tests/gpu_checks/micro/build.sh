#!/bin/bash
# builds the GEMM micro-benchmark binaries next to their sources (run from the repo root)
set -e
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I transformer-pointer-critic_b200/csrc"
nvcc $F -DGEMM_PROFILE tests/gpu_checks/micro/gemm_phase.cu -lcuda -o tests/gpu_checks/micro/gemm_phase 2>&1 | grep -v "warning\|^$\|declared but never\|Remark\|\^" || true
nvcc $F tests/gpu_checks/micro/gemm_phase.cu -lcuda -o tests/gpu_checks/micro/gemm_plain 2>&1 | grep -v "warning\|^$\|declared but never\|Remark\|\^" || true
