for c in 1 0; do
FRB_MULTI=2 FRB_MULTI_COOP=$c timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_c$c.log 2>&1 || tail -5 gpurun_out/r2c_bench_c$c.log
tail -1 gpurun_out/r2c_bench_c$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH coop=$c', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'])"
done
FRB_MULTI=2 python tools/diag_multi.py m2c ir_50 8 2>&1 | tail -1
FRB_MULTI=0 python tools/diag_multi.py ref ir_50 8 2>&1 | tail -1
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag_ref.npy")[0]; x=np.load("gpurun_out/diag_m2c.npy")[0]
print("coop bit-identical:", bool(np.array_equal(r,x)))
P
