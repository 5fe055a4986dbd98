# 2-GPU records of the final build (r02g): multi-GPU correctness tests, strong scaling at global batch 1024, weak scaling
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q > $O/r02g_gpu_multi_tests.log 2>&1
tail -3 $O/r02g_gpu_multi_tests.log
timeout 200 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --steps 4 --warmup 3 --scaling strong --global-batch 1024 --no-cpu > $O/r02g_strong_2gpu_gb1024.json 2> $O/r02g_strong_2gpu.err
timeout 200 $TR --nproc-per-node 2 --master-port 29572 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > $O/r02g_weak_2gpu_b1024.json 2> $O/r02g_weak_2gpu.err
for f in $O/r02g_strong_2gpu_gb1024.json $O/r02g_weak_2gpu_b1024.json; do cut -c1-260 $f; done
