set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_ladder.py conv > gpurun_out/r1c_ladder_conv.log 2>&1; tail -3 gpurun_out/r1c_ladder_conv.log
for cfg in "2 1" "2 0" "1 0"; do
  set -- $cfg
  FRB_CONV_MODE=$1 FRB_SLAB=$2 timeout 300 python tools/microbench_gemm.py conv > gpurun_out/r1c_mb_mode$1_slab$2.log 2>&1
  cat gpurun_out/r1c_mb_mode$1_slab$2.log
  FRB_CONV_MODE=$1 FRB_SLAB=$2 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_mode$1_slab$2.log 2>&1
  tail -1 gpurun_out/r1c_bench_mode$1_slab$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks'])"
done
