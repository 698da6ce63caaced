FRB_MULTI=0 python tools/diag_multi.py ref ir_50 8 2>&1 | tail -1
python tools/diag_multi.py s2 ir_50 8 2>&1 | tail -1
FRB_MULTI=0 python tools/diag_multi.py ref101 ir_101 256 2>&1 | tail -1
python tools/diag_multi.py s2101 ir_101 256 2>&1 | tail -1
python - <<'P'
import numpy as np
for a,b in [("ref","s2"),("ref101","s2101")]:
    r=np.load(f"gpurun_out/diag_{a}.npy"); x=np.load(f"gpurun_out/diag_{b}.npy")
    print(b, "faces differing per rep:", [int((x[k]!=r[0]).any(1).sum()) for k in range(3)])
P
FRB_MULTI=0 python tools/diag_multi2.py ref 2>&1 | tail -1
python tools/diag_multi2.py s2 2>&1 | tail -1
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag2_ref.npy"); x=np.load("gpurun_out/diag2_s2.npy")
print("B=1024 faces differing per rep:", [int((x[k]!=r[0]).any(1).sum()) for k in range(4)])
P
for m in 0 1 0 1; do
FRB_SLAB_MULTI=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_s$m.log 2>&1 || tail -5 gpurun_out/r2j_bench_s$m.log
tail -1 gpurun_out/r2j_bench_s$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH slab_multi=$m', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'], d['gpu_launches'])"
done
