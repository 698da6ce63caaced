#!/bin/bash
# Round-2 profile recipe (B200_PROFILING.md): plain run first, then the launch list, then one `ncu --set full` capture per
# kernel of interest.  Every ncu pass only after the same command exited 0 without ncu.  Under ncu the persistent runs
# go out without the cooperative attribute automatically (api.cu: frb_ctx_create).
#   gpurun --timeout 1700 -- 'bash tools/run_profile_r2.sh'      -> gpurun_out/r2p_*
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4"
$B > gpurun_out/r2p_plain.log 2>&1 || { tail -20 gpurun_out/r2p_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_ll.log 2>&1
echo "launch-list rc=$?"
F="--set full --clock-control none --import-source on"
ncu $F -k regex:gemm2_multi_sm100_kernel -s 4 -c 1 -o gpurun_out/r2p_multi $B > gpurun_out/r2p_ncu_multi.log 2>&1
ncu $F -k regex:match_filter2_kernel -s 4 -c 1 -o gpurun_out/r2p_match256 $B > gpurun_out/r2p_ncu_match256.log 2>&1
ncu $F -k regex:match_filter2_kernel -s 12 -c 1 -o gpurun_out/r2p_match4096 $B > gpurun_out/r2p_ncu_match4096.log 2>&1
ncu $F -k regex:match_finalize_kernel -s 4 -c 1 -o gpurun_out/r2p_finalize $B > gpurun_out/r2p_ncu_finalize.log 2>&1
ncu $F -k regex:stem_tc_kernel -s 4 -c 1 -o gpurun_out/r2p_stem $B > gpurun_out/r2p_ncu_stem.log 2>&1
ncu $F -k regex:conv_slab_sm100_kernel -s 60 -c 1 -o gpurun_out/r2p_slab112 $B > gpurun_out/r2p_ncu_slab112.log 2>&1
ncu $F -k regex:conv_slab_sm100_kernel -s 70 -c 1 -o gpurun_out/r2p_slab28 $B > gpurun_out/r2p_ncu_slab28.log 2>&1
# one shard of the 8-rank sharded match (4096 probes x 125 k rows) on its own
FRB_N=125000 python tools/bench_match.py 4096 > gpurun_out/r2p_match125k_plain.log 2>&1 && \
FRB_N=125000 ncu $F -k regex:match_filter2_kernel -s 6 -c 1 -o gpurun_out/r2p_match125k python tools/bench_match.py 4096 > gpurun_out/r2p_ncu_match125k.log 2>&1
# gpurun brings back at most 64 MiB: export the pages that get read and drop the reports (9 MB each)
for f in gpurun_out/r2p_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  case $b in *slab112|*match125k) ncu -i $f --page source --csv > ${b}_source.csv 2>/dev/null;; esac
  rm -f $f
done
ls -la gpurun_out/r2p_*
