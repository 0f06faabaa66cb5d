for v in 0 1; do
B200RT_WHITTED_PERSISTENT=$v ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_whitted_$v.csv python tools/run_whitted.py opaque > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/launches_whitted_$v.csv | head -12
done
