FRB_MULTI=0 python tools/diag_multi.py ref ir_50 8 2>&1 | tail -1
FRB_MULTI=1 python tools/diag_multi.py m1 ir_50 8 2>&1 | tail -1
FRB_MULTI=1 FRB_PDL=0 python tools/diag_multi.py m1nopdl ir_50 8 2>&1 | tail -1
FRB_MULTI=1 FRB_MULTI_MAXRUN=2 python tools/diag_multi.py m1run2 ir_50 8 2>&1 | tail -1
FRB_MULTI=1 FRB_MULTI_MAXRUN=6 python tools/diag_multi.py m1run6 ir_50 8 2>&1 | tail -1
FRB_MULTI=1 FRB_MULTI_DEBUG=1 python tools/diag_multi.py m1sleep ir_50 8 2>&1 | tail -1
FRB_MULTI=1 FRB_MULTI_DEBUG=2 python tools/diag_multi.py m1nores ir_50 8 2>&1 | tail -1
python - <<'P'
import numpy as np
ref=np.load("gpurun_out/diag_ref.npy")[0]
for t in ["m1","m1nopdl","m1run2","m1run6","m1sleep","m1nores"]:
    try:
        x=np.load(f"gpurun_out/diag_{t}.npy")[0]
        cos=(x*ref).sum(1)/np.linalg.norm(x,axis=1)/np.linalg.norm(ref,axis=1)
        print(t, "cos vs per-layer launches:", np.round(cos,5).tolist())
    except Exception as e: print(t, e)
P
