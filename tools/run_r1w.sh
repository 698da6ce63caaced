# persistent multi-layer runs (FRB_MULTI): correctness, then bench A/B
set -x
FRB_MULTI=1 timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -8
for m in 0 1 0 1; do
FRB_MULTI=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r1w_bench_m$m.log 2>&1 || tail -5 gpurun_out/r1w_bench_m$m.log
tail -1 gpurun_out/r1w_bench_m$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH multi=$m', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['roofline']['kernel'][:30], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['gpu_launches'])"
done
