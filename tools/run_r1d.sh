set -x
export FRB_SLAB=0
for shp in "256 56 64 64" "256 28 128 128" "256 14 256 256"; do
  tag=$(echo $shp | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:gemm2_sm100 -s 5 -c 1 -o gpurun_out/r1d_$tag python tools/microbench_gemm.py one $shp > gpurun_out/r1d_ncu_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
