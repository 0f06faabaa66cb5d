for t in "tests/test_gpu_parity.py -k whitted" "tests/test_gpu_optix_parity.py -k whitted"; do
  echo "=== $t"
  timeout 120 python -m pytest $t -x -q > /tmp/o.txt 2>&1; echo "rc=$?"; tail -12 /tmp/o.txt | cut -c1-300
done
timeout 120 python tools/time_small.py 2>&1 | tail -1
