export FRB_SLAB=1
for shp in "32 56 64 64"; do
  tag=$(echo $shp | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:conv_slab -s 5 -c 1 -o gpurun_out/r1j_slab_$tag python tools/microbench_gemm.py one $shp > gpurun_out/r1j_ncu_$tag.log 2>&1
done
