# full GPU suite + default bench (with cpu baseline) + reference arm
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1p_pytest.log 2>&1; tail -15 gpurun_out/r1p_pytest.log
timeout 600 python bench.py > gpurun_out/r1p_bench.log 2>&1; tail -1 gpurun_out/r1p_bench.log | cut -c1-3000
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/r1p_bench60.log 2>&1; tail -1 gpurun_out/r1p_bench60.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH60', d['value'], d['embed_ms'], d['match_ms'], d['clocks'], d['e2e'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1p_ref.log 2>&1; tail -1 gpurun_out/r1p_ref.log | cut -c1-800
