# two-stream fork of the encoder stacks in the rollout: equality tests, GPU suite, A / B
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_rollout_graph.py -x -q > $O/r02g_fork_test.log 2>&1
tail -8 $O/r02g_fork_test.log
timeout 500 python -m pytest tests -m gpu -x -q > $O/r02g_gpu_tests4.log 2>&1
tail -3 $O/r02g_gpu_tests4.log
for g in 0 2; do
  TPC_ROLLOUT_FORK=$g timeout 200 python bench.py --batch 128 --nodes 10 --rules 20 --steps 20 --warmup 5 --no-cpu > $O/r02g_cfg1_fork$g.json 2>> $O/r02g_fork.err
  TPC_ROLLOUT_FORK=$g timeout 200 python bench.py --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r02g_b128_fork$g.json 2>> $O/r02g_fork.err
  TPC_ROLLOUT_FORK=$g timeout 200 python tests/gpu_checks/bench_tester.py > $O/r02g_tester_fork$g.json 2>> $O/r02g_fork.err
done
for f in $O/r02g_cfg1_fork?.json $O/r02g_b128_fork?.json $O/r02g_tester_fork?.json; do cut -c1-230 $f; done
tail -5 $O/r02g_fork.err
