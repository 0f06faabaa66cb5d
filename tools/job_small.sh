for w in duck_raycast whitted_duck playground; do
  for impl in reference b200rt; do
    python bench.py --workload $w --impl $impl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_${w}_${impl}.json 2> gpurun_out/r02b_${w}_${impl}.err
    echo "$w $impl rc=$?"
  done
done
python -m pytest tests -m gpu -q -rs 2>&1 | grep -i skip
