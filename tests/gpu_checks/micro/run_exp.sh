cd tests/gpu_checks/micro
echo "fwd FFN1 no mask : $(./gemm_plain 1546240 128 1 0 512 0 | head -1)"
echo "fwd FFN1 mask out: $(./gemm_plain 1546240 128 1 0 512 1 | head -1)"
echo "no gate          : $(./gemm_plain 1546240 128 0 0 512 0 | head -1)"
echo "gate by aux h    : $(./gemm_plain 1546240 128 0 0 512 3 | head -1)"
echo "gate by bit image: $(./gemm_plain 1546240 128 0 0 512 2 | head -1)"
