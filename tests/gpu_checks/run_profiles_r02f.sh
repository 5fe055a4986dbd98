set -x
python tests/gpu_checks/roofline_probe.py > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r02f_gemm_tc_ffn2 -f python tests/gpu_checks/roofline_probe.py > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02f_launches_raw.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out/
