#!/bin/bash
# Late round-2 captures: the conv2 (shortcut) Cout = 64 slab layer with the TMA shortcut, and the 6-CTA stem.
#   gpurun --timeout 900 -- 'bash tools/run_profile_r2_late.sh'   -> gpurun_out/r2q_*_raw.csv
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4"
F="--set full --clock-control none --import-source on"
$B > gpurun_out/r2q_plain.log 2>&1 || { tail -20 gpurun_out/r2q_plain.log; exit 1; }
# 30 conv_slab launches per embed (5 + 1 + 24); the third and fifth are conv_slab_sm100_kernel<64,1,true> (conv2 of units 2, 3)
ncu $F -k regex:conv_slab_sm100_kernel -s 62 -c 1 -o gpurun_out/r2q_slab56res $B > gpurun_out/r2q_ncu_slab56res.log 2>&1
ncu $F -k regex:stem_tc_kernel -s 4 -c 1 -o gpurun_out/r2q_stem $B > gpurun_out/r2q_ncu_stem.log 2>&1
for f in gpurun_out/r2q_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  rm -f $f
done
ls -la gpurun_out/r2q_*
