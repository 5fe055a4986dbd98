# fused LayerNorm-backward GEMM epilogue: unit check + timing, GPU tests, A / B of the bench step
set -x
O=gpurun_out
timeout 200 python -c "from tests.gpu_checks import check_gemm; check_gemm.check_gemm_ln_bwd(timing=True)" > $O/r02g_lnb_check.log 2>&1
tail -12 $O/r02g_lnb_check.log
timeout 400 python -m pytest tests -m gpu -x -q > $O/r02g_gpu_tests.log 2>&1
tail -3 $O/r02g_gpu_tests.log
TPC_FUSE_LNB=0 timeout 200 python bench.py --no-cpu > $O/r02g_bench_unfused.json 2> $O/r02g_unf.err
timeout 200 python bench.py --no-cpu > $O/r02g_bench_fused.json 2> $O/r02g_f.err
TPC_FUSE_LNB=0 timeout 200 python bench.py --no-cpu > $O/r02g_bench_unfused2.json 2>> $O/r02g_unf.err
timeout 200 python bench.py --no-cpu > $O/r02g_bench_fused2.json 2>> $O/r02g_f.err
for f in $O/r02g_bench_unfused.json $O/r02g_bench_fused.json $O/r02g_bench_unfused2.json $O/r02g_bench_fused2.json; do cut -c1-160 $f; done
