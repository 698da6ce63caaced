FRB_MULTI=0 python tools/diag_multi2.py ref 2>&1 | tail -1
for i in 1 2 3; do python tools/diag_multi2.py m2_$i 2>&1 | tail -1; done
FRB_MULTI=1 python tools/diag_multi2.py m1 2>&1 | tail -1
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag2_ref.npy")
for t in ["m2_1","m2_2","m2_3","m1"]:
    x=np.load(f"gpurun_out/diag2_{t}.npy")
    for k in range(4):
        bad=np.nonzero((x[k]!=r[0]).any(1))[0]
        print(t, "rep",k, "faces differing from per-layer launches:", len(bad), bad[:20].tolist())
P
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
