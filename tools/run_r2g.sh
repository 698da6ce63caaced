FRB_MULTI=0 python tools/diag_multi2.py ref 2>&1 | tail -1
FRB_MULTI_DEBUG=4 python tools/diag_multi2.py norot 2>&1 | tail -1
FRB_MULTI_DEBUG=8 python tools/diag_multi2.py allfence 2>&1 | tail -1
FRB_MULTI_DEBUG=16 python tools/diag_multi2.py waitmore 2>&1 | tail -1
FRB_MULTI_DEBUG=12 python tools/diag_multi2.py norot_allfence 2>&1 | tail -1
python - <<'P'
import numpy as np
r=np.load("gpurun_out/diag2_ref.npy")
for t in ["norot","allfence","waitmore","norot_allfence"]:
    x=np.load(f"gpurun_out/diag2_{t}.npy")
    print(t, [int((x[k]!=r[0]).any(1).sum()) for k in range(4)])
P
