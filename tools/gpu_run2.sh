timeout 600 python -m pytest tests -x -q -m gpu > /tmp/o.txt 2>&1; echo "pytest rc=$?"; tail -5 /tmp/o.txt | cut -c1-400
timeout 200 python tools/time_small.py 2>&1 | tail -1
