# probes of the r02g build: attention kernels at rollout-sized vs update-sized state counts, ncu --set full of the
# LayerNorm-backward GEMM, memcheck of the smoke path
set -x
O=gpurun_out
for s in 1024 10240; do STATES=$s SKIP_CHECK=1 timeout 120 python tests/gpu_checks/bench_attention.py; done > $O/r02g_attention_states.log 2>&1
cat $O/r02g_attention_states.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 12 -c 1 -o $O/r02g_gemm_ln_bwd -f python -c "from tests.gpu_checks import check_gemm; check_gemm.check_gemm_ln_bwd(timing=True)" > $O/r02g_ncu_lnb.log 2>&1
tail -3 $O/r02g_ncu_lnb.log
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02g_memcheck.log 2>&1
echo memcheck rc=$?
tail -6 $O/r02g_memcheck.log
