# Round-1 profile D: the persistent multi-layer run (default schedule).  ncu passes only after the plain run exited 0.
# ncu cannot replay a cooperative launch of this kernel (LaunchFailed): the captures use FRB_MULTI_COOP=0 (same kernel,
# same grid, no gang-scheduling attribute).
set -x
export FRB_MULTI_COOP=0
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_ll.log 2>&1
tail -2 gpurun_out/r1d_ncu_ll.log
ncu --set full --clock-control none --import-source on -k regex:gemm2_multi_sm100_kernel -s 4 -c 1 -o gpurun_out/r1d_multi python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1d_ncu_multi.log 2>&1
tail -3 gpurun_out/r1d_ncu_multi.log
unset FRB_MULTI_COOP
python bench.py > gpurun_out/r1d_bench.log 2>&1; tail -1 gpurun_out/r1d_bench.log | cut -c1-1200
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1d_ref.log 2>&1; tail -1 gpurun_out/r1d_ref.log | cut -c1-300
ls -la gpurun_out/r1d_*
