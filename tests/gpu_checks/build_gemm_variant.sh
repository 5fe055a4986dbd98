#!/bin/bash
# Timing experiments on gemm_tc.cu: builds transformer-pointer-critic_b200/libtpc_<name>.so with extra -D switches (results of
# such builds are WRONG by design; they only answer "what does this part of the pipeline cost").
#   tests/gpu_checks/build_gemm_variant.sh name -DGEMM_DBG_MMA4 ...
# then on the GPU box: TPC_LIB=/root/repo/transformer-pointer-critic_b200/libtpc_<name>.so python tests/gpu_checks/bench_kernels.py
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
pkg=transformer-pointer-critic_b200
python -m tpc_b200.build >/dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr -I include "$@" -c $pkg/csrc/gemm_tc.cu -o /tmp/gemm_tc_$name.o
objs=$(ls $pkg/csrc/build/*.o | grep -v gemm_tc.o)
/usr/local/cuda/bin/nvcc -shared -o $pkg/libtpc_$name.so $objs /tmp/gemm_tc_$name.o -gencode arch=compute_100a,code=sm_100a
echo $pkg/libtpc_$name.so
