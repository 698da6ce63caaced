# Round-1 profile C (final state of the round): launch list + full captures of the three dominant kernels.
# Every ncu pass runs only after the same command exited 0 without ncu.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_sm100_kernel -s 234 -c 1 -o gpurun_out/r1c_gemm2_256 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_gemm2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_slab_sm100_kernel -s 100 -c 1 -o gpurun_out/r1c_slab_128 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_slab.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_filter2_kernel -s 3 -c 1 -o gpurun_out/r1c_match2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_match.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_finalize_kernel -s 3 -c 1 -o gpurun_out/r1c_finalize python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_fin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 3 -c 1 -o gpurun_out/r1c_stem python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_stem.log 2>&1
ls -la gpurun_out/r1c_*
