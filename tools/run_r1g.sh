for b in 256 1024; do
FRB_SLAB=1 FRB_SLAB_N256=1 timeout 100 python tools/microbench_gemm.py one $b 28 128 256
FRB_SLAB=0 timeout 100 python tools/microbench_gemm.py one $b 28 128 256
done
