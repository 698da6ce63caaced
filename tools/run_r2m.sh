for w in 0 56 0 56; do
FRB_SLAB_MINW=$w timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_w$w.log 2>&1 || tail -5 gpurun_out/r2m_bench_w$w.log
tail -1 gpurun_out/r2m_bench_w$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH slab_min_w=$w', d['value'], d['embed_ms'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
FRB_SLAB_MINW=56 python tools/diag_multi.py w56 ir_101 64 2>&1 | tail -1
FRB_MULTI=0 python tools/diag_multi.py ref64 ir_101 64 2>&1 | tail -1
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag_ref64.npy")[0]; x=np.load("gpurun_out/diag_w56.npy")[0]
c=(r*x).sum(1)/np.linalg.norm(r,axis=1)/np.linalg.norm(x,axis=1)
print("im2col-vs-slab stage 2: min cosine", c.min())
P
