# ncu --set full of the weight-resident GEMM (forward FFN1 128 -> 512, split mode, relu) and the weight gradient, final build
set -x
python tests/gpu_checks/profile_kernels.py > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bres -s 1 -c 1 -o gpurun_out/r02f_gemm_bres_ffn1 -f python tests/gpu_checks/profile_kernels.py > gpurun_out/ncu_bres.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 1 -c 1 -o gpurun_out/r02f_wgrad -f python tests/gpu_checks/profile_kernels.py > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
