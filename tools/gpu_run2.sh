timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_instance" > /tmp/o.txt 2>&1; echo "pytest rc=$?"; tail -15 /tmp/o.txt | cut -c1-400
