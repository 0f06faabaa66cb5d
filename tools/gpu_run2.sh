timeout 300 python -m pytest tests -x -q -m gpu -k "shim" > /tmp/o.txt 2>&1; echo "rc=$?"; tail -30 /tmp/o.txt | cut -c1-400
