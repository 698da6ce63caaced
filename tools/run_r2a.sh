FRB_MULTI=0 python tools/diag_multi.py ref ir_50 8 2>&1 | tail -1
FRB_MULTI=2 python tools/diag_multi.py m2 ir_50 8 2>&1 | tail -1
FRB_MULTI=0 python tools/diag_multi.py ref101 ir_101 256 2>&1 | tail -1
FRB_MULTI=2 python tools/diag_multi.py m2101 ir_101 256 2>&1 | tail -1
python - <<'P'
import numpy as np
for a,b in [("ref","m2"),("ref101","m2101")]:
    r=np.load(f"gpurun_out/diag_{a}.npy")[0]; x=np.load(f"gpurun_out/diag_{b}.npy")[0]
    print(b, "bit-identical to per-layer launches:", bool(np.array_equal(r,x)), float(np.abs(r-x).max()))
P
FRB_MULTI=2 timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py tests/test_gpu_kernels.py tests/test_flows.py -m gpu -x -q 2>&1 | tail -4
for m in 0 2 1 2; do
FRB_MULTI=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_m$m.log 2>&1 || tail -5 gpurun_out/r2a_bench_m$m.log
tail -1 gpurun_out/r2a_bench_m$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH multi=$m', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'], d['gpu_launches'])"
done
echo "MULTI=2"; FRB_MULTI=2 timeout 300 python tools/bench_small.py 2>&1 | tail -5
