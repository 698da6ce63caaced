set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1u_pytest.log 2>&1; tail -4 gpurun_out/r1u_pytest.log
timeout 300 python tools/bench_match.py 256 1024 4096 2>&1 | tail -4
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r1u_bench.log 2>&1
tail -1 gpurun_out/r1u_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'])"
