# split-K tail (FRB_TAIL_SPLIT): correctness at full batch, microbench, bench A/B
set -x
for ts in 0 1; do
FRB_TAIL_SPLIT=$ts timeout 300 python tools/gpu_ladder.py conv_big > gpurun_out/r1r_ladder_big_ts$ts.log 2>&1; grep -c "'ok': True" gpurun_out/r1r_ladder_big_ts$ts.log; grep "'ok': False\|Error\|error\|timeout" gpurun_out/r1r_ladder_big_ts$ts.log | cut -c1-400
done
FRB_QUAD=1 timeout 300 python tools/gpu_ladder.py conv_big > gpurun_out/r1r_ladder_big_quad.log 2>&1; grep -c "'ok': True" gpurun_out/r1r_ladder_big_quad.log
for ts in 0 1; do
echo "TAIL_SPLIT=$ts"
FRB_TAIL_SPLIT=$ts timeout 200 python tools/microbench_gemm.py conv 2>&1 | tee gpurun_out/r1r_mb_ts$ts.log | tail -12
done
FRB_TAIL_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -8
for ts in 0 1 0 1; do
FRB_TAIL_SPLIT=$ts timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r1r_bench_ts$ts.log 2>&1
tail -1 gpurun_out/r1r_bench_ts$ts.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH ts=$ts', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['roofline']['avg_launch_us'], d['roofline']['frac'])"
done
