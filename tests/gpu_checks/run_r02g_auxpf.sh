# L2 prefetch of the next tile's residual boxes in the LayerNorm epilogues: A / B (TPC_AUX_PREFETCH=0/1), then the GPU suite
set -x
O=gpurun_out
for g in 0 1; do
  TPC_AUX_PREFETCH=$g CHUNK=10240 timeout 100 python tests/gpu_checks/bench_kernels.py 2>&1 | grep "LN\|dout1\|dx 384" > $O/r02g_auxpf${g}_kernels.txt
  TPC_AUX_PREFETCH=$g timeout 100 python -c "from tests.gpu_checks import check_gemm; check_gemm.check_gemm_ln_bwd(timing=True)" 2>&1 | grep "K=" > $O/r02g_auxpf${g}_lnb.txt
done
for g in 0 1 0 1; do TPC_AUX_PREFETCH=$g timeout 200 python bench.py --no-cpu --steps 4 2>> $O/r02g_auxpf.err | cut -c1-140; done > $O/r02g_auxpf_bench.txt
cat $O/r02g_auxpf0_kernels.txt $O/r02g_auxpf1_kernels.txt $O/r02g_auxpf0_lnb.txt $O/r02g_auxpf1_lnb.txt $O/r02g_auxpf_bench.txt
timeout 300 python -m pytest tests -m gpu -q > $O/r02g_gpu_tests_final3.log 2>&1; tail -2 $O/r02g_gpu_tests_final3.log
