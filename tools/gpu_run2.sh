for t in "tests/test_gpu_parity.py -k whitted" "tests/test_gpu_optix_parity.py -k whitted" "tests/test_gpu_optix_parity.py -k raycasting"; do
  echo "=== $t"
  timeout 120 python -m pytest $t -x -q > /tmp/o.txt 2>&1; echo "rc=$?"; tail -25 /tmp/o.txt
done
