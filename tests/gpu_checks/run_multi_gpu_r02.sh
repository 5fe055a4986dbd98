set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
$TR --nproc-per-node 2 --master-port 29540 bench.py --gpus 2 --check --batch 24 --nodes 10 --rules 20 > $O/r02_check_2gpu_10x20.json 2> $O/r02_check_2gpu.err
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --check --batch 8 --nodes 50 --rules 100 > $O/r02_check_8gpu_50x100.json 2> $O/r02_check_8gpu.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2955$n bench.py --gpus $n --steps 4 --warmup 3 --scaling strong --global-batch 1024 --no-cpu > $O/r02_strong_${n}gpu_gb1024.json 2> $O/r02_strong_${n}gpu.err
done
$TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --global-batch 4096 --reward global_dominant --no-cpu > $O/r02_cfg3_fair_8gpu_gb4096.json 2> $O/r02_cfg3_fair.err
$TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --global-batch 4096 --reward reduced_node_usage --no-cpu > $O/r02_cfg3_cost_8gpu_gb4096.json 2> $O/r02_cfg3_cost.err
$TR --nproc-per-node 8 --master-port 29563 tests/gpu_checks/bench_cvrp.py 256 16384 > $O/r02_cvrp_sweep_8gpu.jsonl 2> $O/r02_cvrp_8gpu.err
python -m pytest tests/test_gpu_multi.py -q > $O/r02_gpu_multi_test.log 2>&1
tail -3 $O/r02_gpu_multi_test.log
for f in $O/r02_check_*json $O/r02_strong_*json $O/r02_cfg3_*json $O/r02_cvrp_sweep_8gpu.jsonl; do echo $f; cut -c1-330 $f; done
