# 1-GPU records of the final round-2 build (r02g: LayerNorm-backward GEMM epilogue, rollout graph + two-stream fork)
set -x
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q > $O/r02g_gpu_tests_final.log 2>&1
tail -3 $O/r02g_gpu_tests_final.log
timeout 300 python bench.py > $O/r02g_bench_b1024.json 2> $O/r02g_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02g_bench_reference_arm.json 2> $O/r02g_ref.err
timeout 200 python bench.py --batch 128 --nodes 10 --rules 20 --steps 20 --warmup 5 --no-cpu > $O/r02g_bench_cfg1_shipped_b128.json 2>> $O/r02g_bench.err
timeout 200 python tests/gpu_checks/bench_tester.py > $O/r02g_tester_cell_50x100_b1024.json 2>> $O/r02g_bench.err
CHUNK=10240 timeout 120 python tests/gpu_checks/bench_kernels.py > $O/r02g_bench_kernels.txt 2>&1
timeout 120 python tests/gpu_checks/check_golden_model.py > $O/r02g_golden_model_check.log 2>&1
for f in $O/r02g_bench_b1024.json $O/r02g_bench_reference_arm.json $O/r02g_bench_cfg1_shipped_b128.json $O/r02g_tester_cell_50x100_b1024.json; do cut -c1-260 $f; done
tail -20 $O/r02g_bench_kernels.txt
# ncu: launch list of one step, --set full of the LayerNorm-backward GEMM (K = 512, 1.55 M rows)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $O/r02g_launches_raw.csv python bench.py --steps 1 --warmup 0 --no-cpu > $O/r02g_ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel.1 -s 8 -c 1 -o $O/r02g_gemm_tc_ln_bwd -f python -c "from tests.gpu_checks import check_gemm; check_gemm.check_gemm_ln_bwd(timing=True)" > $O/r02g_ncu_lnb2.log 2>&1
ls -la $O/r02g_launches_raw.csv $O/r02g_gemm_tc_ln_bwd.ncu-rep
