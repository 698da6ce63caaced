python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ll_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ll_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ll_ncu.log 2>&1
tail -3 gpurun_out/ll_ncu.log
