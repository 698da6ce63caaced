for d in 0 512 0 512; do
FRB_SLAB_DEBUG=$d timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_d$d.log 2>&1 || tail -5 gpurun_out/r2k_bench_d$d.log
tail -1 gpurun_out/r2k_bench_d$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH slab_debug=$d', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'])"
done
FRB_SLAB_DEBUG=512 timeout 100 python tools/microbench_gemm.py slab28 2>&1 | tail -1
FRB_SLAB_DEBUG=0 timeout 100 python tools/microbench_gemm.py slab28 2>&1 | tail -1
