cd tests/gpu_checks/micro
./gemm_phase 1546240 128 1 0 512 | head -16
./gemm_phase 1546240 512 1 1 128 | head -12
