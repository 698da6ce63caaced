set -x
python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench60.log 2>&1
tail -1 gpurun_out/bench60.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['embed_ms'], d['match_ms'], d['clocks'])"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 363 -c 1 -o gpurun_out/prof_stage3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 302 -c 1 -o gpurun_out/prof_stage1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 320 -c 1 -o gpurun_out/prof_stage2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s2.log 2>&1
ls -la gpurun_out/*.ncu-rep
