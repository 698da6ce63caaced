timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py tests/test_flows.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_$i.log 2>&1 || tail -5 gpurun_out/r2i_bench_$i.log
tail -1 gpurun_out/r2i_bench_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'])"
done
