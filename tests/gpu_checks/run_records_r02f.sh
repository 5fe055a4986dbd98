# 1-GPU records of the final round-2 build (FP16 split forward GEMMs, two-warpgroup epilogues)
set -x
O=gpurun_out
timeout 300 python bench.py --batch 128 --nodes 10 --rules 20 --steps 20 --warmup 5 --no-cpu > $O/r02f_bench_cfg1_shipped_b128.json 2> $O/cfg1.err
timeout 300 python tests/gpu_checks/bench_tester.py > $O/r02f_tester_cell_50x100_b1024.json 2> $O/tester.err
timeout 300 python tests/gpu_checks/bench_cvrp.py > $O/r02f_cvrp_rollout_sweep_1gpu.jsonl 2> $O/cvrp.err
CHUNK=10240 timeout 120 python tests/gpu_checks/bench_kernels.py > $O/r02f_bench_kernels.txt 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02f_bench_reference_arm.json 2> $O/ref.err
for f in $O/r02f_bench_cfg1_shipped_b128.json $O/r02f_tester_cell_50x100_b1024.json $O/r02f_bench_reference_arm.json; do cut -c1-300 $f; done
cat $O/r02f_bench_kernels.txt; tail -4 $O/r02f_cvrp_rollout_sweep_1gpu.jsonl | cut -c1-200
