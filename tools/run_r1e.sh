set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_ladder.py conv > gpurun_out/r1e_ladder_conv.log 2>&1; grep -c "'ok': True" gpurun_out/r1e_ladder_conv.log; grep "'ok': False" gpurun_out/r1e_ladder_conv.log | cut -c1-300; tail -2 gpurun_out/r1e_ladder_conv.log | cut -c1-300
FRB_SLAB=1 timeout 300 python tools/microbench_gemm.py conv 2>&1 | tee gpurun_out/r1e_mb_slab1.log
for shp in "256 112 64 64" "256 56 64 128" "256 28 128 256"; do FRB_SLAB=1 timeout 100 python tools/microbench_gemm.py one $shp; FRB_SLAB=0 timeout 100 python tools/microbench_gemm.py one $shp; done 2>&1 | tee gpurun_out/r1e_mb_more.log
timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -5
for sl in 1 0; do
FRB_SLAB=$sl timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_slab$sl.log 2>&1
tail -1 gpurun_out/r1e_bench_slab$sl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks'])"
done
