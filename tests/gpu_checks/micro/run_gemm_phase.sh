cd tests/gpu_checks/micro
for cfg in "1546240 512 1 1 128" "1546240 512 0 0 128" "1546240 128 1 1 128" "1546240 384 0 0 128" "1546240 128 1 0 512" "1546240 128 1 0 384" "1546240 128 0 0 512"; do
  echo "== plain $cfg"; ./gemm_plain $cfg | head -1
  echo "== phase $cfg"; ./gemm_phase $cfg
done
