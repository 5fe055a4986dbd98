# (1) bench.py under ncu must survive the second rollout (graph capture is skipped under an injection tool);
# (2) ncu --set full of the LayerNorm-backward GEMM (the 14th gemm_tc_kernel launch of the check is a fused K = 512 one)
set -x
O=gpurun_out
env | grep -i "inject\|NV_COMPUTE" 
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02g_ncu_small.csv python bench.py --batch 16 --nodes 5 --rules 7 --steps 2 --warmup 1 --no-cpu > $O/r02g_ncu_small.log 2>&1
echo rc=$?; grep -c . $O/r02g_ncu_small.csv; grep "ERROR" $O/r02g_ncu_small.csv | head -3; cut -c1-200 $O/r02g_ncu_small.log | tail -3
timeout 100 ncu --metrics gpu__time_duration.sum -c 5 python -c "import os; print({k: v for k, v in os.environ.items() if 'INJECT' in k or 'NV_COMPUTE' in k or 'NSIGHT' in k.upper()})" 2>&1 | tail -3
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 22 -c 1 -o $O/r02g_gemm_tc_ln_bwd -f python -c "from tests.gpu_checks import check_gemm; check_gemm.check_gemm_ln_bwd(timing=True)" > $O/r02g_ncu_lnb2.log 2>&1
ls -la $O/r02g_gemm_tc_ln_bwd.ncu-rep
