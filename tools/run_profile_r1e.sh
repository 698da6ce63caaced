# Profile E: the stage-1 slab launches (112x112 and 56x56, Cout 64) and the strided pair kernel, full captures
set -x
export FRB_MULTI_COOP=0
ncu --set full --clock-control none --import-source on -k regex:conv_slab_sm100_kernel -s 15 -c 2 -o gpurun_out/r1e_slab64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_ncu_slab64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_sm100_kernel -s 6 -c 1 -o gpurun_out/r1e_gemm2_64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_ncu_gemm2_64.log 2>&1
ls -la gpurun_out/r1e_*
