export FRB_SLAB=1
for d in 0 16; do echo "debug=$d"; for shp in "256 112 64 64" "256 56 64 64" "256 28 128 128"; do FRB_SLAB_DEBUG=$d timeout 100 python tools/microbench_gemm.py one $shp; done; done
FRB_SLAB_DEBUG=16 timeout 600 python tools/gpu_ladder.py conv > gpurun_out/r1i_ladder_conv.log 2>&1; grep -c "'ok': True" gpurun_out/r1i_ladder_conv.log; grep "'ok': False" gpurun_out/r1i_ladder_conv.log | cut -c1-300
