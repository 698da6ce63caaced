# Round-1 profile B: launch list + full captures of the two dominant kernels (each only after the plain run exits 0)
set -x
timeout 600 python -m pytest tests/test_gpu_match.py -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_bench.log 2>&1 || exit 1
tail -1 gpurun_out/r1b_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks'])"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_sm100_kernel -s 234 -c 1 -o gpurun_out/r1b_gemm2_256 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_gemm2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_slab_sm100_kernel -s 100 -c 1 -o gpurun_out/r1b_slab_128 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_slab.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_filter_kernel -s 3 -c 1 -o gpurun_out/r1b_match python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_match.log 2>&1
ls -la gpurun_out/r1b_*
