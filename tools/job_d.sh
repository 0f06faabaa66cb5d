( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['build']['ms'], d['gpu_launches'], d['clocks'])"
( time python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/bench_default_ref.json').read().strip().splitlines()[-1]); print(d['impl'], d['reference_class'], d['value'], d['e2e']['value'])"
