# quad-cluster weight multicast (FRB_QUAD): correctness ladder, microbench, bench A/B
set -x
FRB_QUAD=1 timeout 300 python tools/gpu_ladder.py conv > gpurun_out/r1q_ladder_conv.log 2>&1; grep -c "'ok': True" gpurun_out/r1q_ladder_conv.log; grep "'ok': False" gpurun_out/r1q_ladder_conv.log | cut -c1-300; tail -3 gpurun_out/r1q_ladder_conv.log | cut -c1-400
for q in 0 1 2; do
echo "QUAD=$q"
FRB_QUAD=$q timeout 200 python tools/microbench_gemm.py conv 2>&1 | tee gpurun_out/r1q_mb_quad$q.log | tail -12
FRB_QUAD=$q timeout 100 python tools/microbench_gemm.py one 256 28 128 256 2>&1 | tail -1
done
FRB_QUAD=1 timeout 600 python -m pytest tests/test_gpu_embed.py tests/test_gpu_kernels.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -5
for q in 0 1 2 0 1; do
FRB_QUAD=$q timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r1q_bench_quad$q.log 2>&1
tail -1 gpurun_out/r1q_bench_quad$q.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH quad=$q', d['value'], d['embed_ms'], d['match_ms'], d['clocks'], d['roofline']['avg_launch_us'])"
done
