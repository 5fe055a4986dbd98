# last check of the final tree: smoke(), GPU suite, one bench line
set -x
O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02g_smoke.log 2>&1; tail -2 $O/r02g_smoke.log
timeout 500 python -m pytest tests -m gpu -q > $O/r02g_gpu_tests_final2.log 2>&1; tail -2 $O/r02g_gpu_tests_final2.log
timeout 200 python bench.py --no-cpu > $O/r02g_bench_b1024_final2.json 2> $O/r02g_bench2.err; cut -c1-200 $O/r02g_bench_b1024_final2.json
