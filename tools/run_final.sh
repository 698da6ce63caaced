set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['ms_per_step'], d['embed_ms'], d['match_ms'], d['clocks'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['backbone_section'], d.get('cpu_baseline'))"
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-250
