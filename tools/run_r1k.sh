mkdir -p gpurun_out
timeout 600 python tools/gpu_ladder.py conv > gpurun_out/r1k_ladder_conv.log 2>&1; grep -c "'ok': True" gpurun_out/r1k_ladder_conv.log; grep "'ok': False" gpurun_out/r1k_ladder_conv.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_kernels.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -3
for pdl in 1 0; do
FRB_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r1k_bench_pdl$pdl.log 2>&1
tail -1 gpurun_out/r1k_bench_pdl$pdl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH pdl=$pdl', d['value'], d['embed_ms'], d['match_ms'], d['clocks'])"
done
