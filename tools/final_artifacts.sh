#!/bin/bash
# Round artefacts in one gpurun call (1 GPU): GPU tests, the bench line and its reference arm, the Cornell line, the ncu launch list of
# the bench command.  Outputs under gpurun_out/; copy what is to be judged into profiles/.
R=${1:-r01}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -2 gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_${R}_synth.json 2> gpurun_out/bench_${R}_synth.err; tail -c 1500 gpurun_out/bench_${R}_synth.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_ref.json 2> gpurun_out/bench_${R}_ref.err
python bench.py --workload cornell --width 768 --height 768 --sample-groups 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_cornell.json 2> gpurun_out/bench_${R}_cornell.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${R}.log 2>&1
echo done
