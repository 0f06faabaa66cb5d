for g in 16; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --sample-groups $g > gpurun_out/bench_r02_n8_g$g.json 2> gpurun_out/bench_r02_n8_g$g.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n8_g$g.json').read().strip().splitlines()[-1]); print('N=8 groups $g', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 4 --warmup 3 --sample-groups 16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=4 groups 16', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python bench.py --steps 3 --warmup 3 --sample-groups 16 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 groups 16', d['value'], d['ms_per_step'], d['e2e']['value'])"
