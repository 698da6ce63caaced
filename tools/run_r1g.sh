export FRB_SLAB=1
for b in 64 128 256 512; do for shp in "$b 56 64 64" "$b 28 128 128"; do FRB_SLAB_DEBUG=0 timeout 100 python tools/microbench_gemm.py one $shp; done; done
