# 8-GPU records of the final build: weak (batch 1024 / GPU) and strong (global batch 1024) scaling
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
$TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 4 --warmup 3 --scaling strong --global-batch 1024 --no-cpu > $O/r02f_strong_8gpu_gb1024.json 2> $O/r02f_strong_8gpu.err
$TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > $O/r02f_weak_8gpu_b1024.json 2> $O/r02f_weak_8gpu.err
for f in $O/r02f_strong_8gpu_gb1024.json $O/r02f_weak_8gpu_b1024.json; do cut -c1-260 $f; done
