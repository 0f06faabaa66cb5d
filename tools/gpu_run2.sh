timeout 300 python -m pytest tests -x -q -m gpu > /tmp/o.txt 2>&1; echo "rc=$?"; tail -12 /tmp/o.txt | cut -c1-300
timeout 120 python tools/time_small.py 2>&1 | tail -1
