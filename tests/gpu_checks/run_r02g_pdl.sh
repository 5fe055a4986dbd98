# programmatic dependent launch on every kernel of the rollout / update loops: GPU suite, then A / B
set -x
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/r02g_gpu_tests3.log 2>&1
tail -4 $O/r02g_gpu_tests3.log
for g in 0 1; do
  TPC_PDL=$g timeout 200 python bench.py --batch 128 --nodes 10 --rules 20 --steps 20 --warmup 5 --no-cpu > $O/r02g_cfg1_pdl$g.json 2>> $O/r02g_pdl.err
  TPC_PDL=$g timeout 200 python bench.py --no-cpu > $O/r02g_bench_pdl$g.json 2>> $O/r02g_pdl.err
done
TPC_PDL=0 timeout 200 python bench.py --no-cpu > $O/r02g_bench_pdl0b.json 2>> $O/r02g_pdl.err
TPC_PDL=1 timeout 200 python bench.py --no-cpu > $O/r02g_bench_pdl1b.json 2>> $O/r02g_pdl.err
for f in $O/r02g_cfg1_pdl?.json $O/r02g_bench_pdl*.json; do cut -c1-200 $f; done
tail -5 $O/r02g_pdl.err
