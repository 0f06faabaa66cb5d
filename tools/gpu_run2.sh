for c in -1 28 35 50 100; do
  echo -n "carveout $c synth: "; B200RT_TRACE_CARVEOUT=$c timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2))"
  echo -n "carveout $c cornell: "; B200RT_TRACE_CARVEOUT=$c timeout 300 python bench.py --workload cornell --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2))"
done
