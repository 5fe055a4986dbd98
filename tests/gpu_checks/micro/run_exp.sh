echo "fwd FFN1 no mask : $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546240 128 1 0 512 0 | head -1)"
echo "fwd FFN1 mask out: $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546240 128 1 0 512 1 | head -1)"
echo "fwd FFN1 mask out ragged: $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546201 128 1 0 512 1 | head -1)"
