for i in 1 2 3; do echo "coop, no pdl (default) #$i"; timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k config4 2>&1 | tail -1; done
for i in 1 2 3; do echo "coop + pdl #$i"; FRB_MULTI_COOP_PDL=1 timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k config4 2>&1 | tail -1; done
for i in 1 2 3; do echo "no coop, pdl #$i"; FRB_MULTI_COOP=0 timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k config4 2>&1 | tail -1; done
for i in 1 2; do echo "per-layer launches #$i"; FRB_MULTI=0 timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k config4 2>&1 | tail -1; done
