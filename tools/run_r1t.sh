set -x
timeout 600 python -m pytest tests/test_gpu_match.py tests/test_identify.py tests/test_gallery_matrix.py tests/test_gpu_aggregate.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_match.py 2>&1 | tail -8
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r1t_bench.log 2>&1
tail -1 gpurun_out/r1t_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'])"
