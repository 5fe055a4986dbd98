cd tests/gpu_checks/micro
for cfg in "1546240 128 1 0 512" "1546240 128 1 0 384"; do
  echo "base:   $(./gemm_plain $cfg | head -1)"
  echo "direct: $(./gemm_direct $cfg | head -1)"
done
./gemm_direct_prof 1546240 128 1 0 512 | head -16
