for g in 148 132 116 100; do
FRB_MULTI_GRID=$g timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_g$g.log 2>&1 || tail -5 gpurun_out/r2d_bench_g$g.log
tail -1 gpurun_out/r2d_bench_g$g.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH grid=$g', d['value'], d['embed_ms'], d['clocks']['sm_mhz'], d['roofline']['avg_launch_us'])"
done
