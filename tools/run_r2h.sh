FRB_MULTI=0 python tools/diag_multi2.py ref 2>&1 | tail -1
for i in 1 2 3 4; do python tools/diag_multi2.py m2_$i 2>&1 | tail -1; done
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag2_ref.npy")
for t in ["m2_1","m2_2","m2_3","m2_4"]:
    x=np.load(f"gpurun_out/diag2_{t}.npy")
    print(t, [int((x[k]!=r[0]).any(1).sum()) for k in range(4)])
P
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_$i.log 2>&1 || tail -5 gpurun_out/r2h_bench_$i.log
tail -1 gpurun_out/r2h_bench_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['embed_ms'], d['match_ms'], d['clocks']['sm_mhz'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])"
done
