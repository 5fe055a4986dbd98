echo "fwd FFN1 mask out: $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546240 128 1 0 512 1 | head -1)"
echo "no gate          : $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546240 128 0 0 512 0 | head -1)"
echo "gate by bit image: $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546240 128 0 0 512 2 | head -1)"
echo "gate, ragged M   : $(timeout 60 tests/gpu_checks/micro/gemm_plain 1546201 128 0 0 512 2 | head -1)"
