timeout 400 python -m pytest tests -x -q -m gpu > /tmp/o.txt 2>&1; echo "rc=$?"; tail -8 /tmp/o.txt | cut -c1-300
timeout 300 python tools/optix_compare.py --skip-synth > gpurun_out/optix_compare6.log 2>&1; echo "compare rc=$?"
