# rollout step loop as a CUDA graph: equality test, GPU suite, A / B at the launch-bound shapes and at the headline
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_rollout_graph.py -x -q > $O/r02g_graph_test.log 2>&1
tail -15 $O/r02g_graph_test.log
timeout 400 python -m pytest tests -m gpu -x -q > $O/r02g_gpu_tests2.log 2>&1
tail -3 $O/r02g_gpu_tests2.log
for g in 0 1; do
  TPC_ROLLOUT_GRAPH=$g timeout 200 python bench.py --batch 128 --nodes 10 --rules 20 --steps 20 --warmup 5 --no-cpu > $O/r02g_cfg1_graph$g.json 2>> $O/r02g_graph.err
  TPC_ROLLOUT_GRAPH=$g timeout 200 python tests/gpu_checks/bench_tester.py > $O/r02g_tester_graph$g.json 2>> $O/r02g_graph.err
  TPC_ROLLOUT_GRAPH=$g timeout 200 python bench.py --no-cpu > $O/r02g_bench_graph$g.json 2>> $O/r02g_graph.err
done
for f in $O/r02g_cfg1_graph?.json $O/r02g_tester_graph?.json $O/r02g_bench_graph?.json; do cut -c1-250 $f; done
tail -5 $O/r02g_graph.err
