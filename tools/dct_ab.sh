#!/bin/bash
# GPU-side A/B of the DCT stage: timing (tools/prof_dct.py) and executed warp instructions (ncu) for each library given
# (default: the in-tree one).  Usage on the box: bash tools/dct_ab.sh [lib ...] ; output in gpurun_out/dct_ab.txt
mkdir -p gpurun_out
: > gpurun_out/dct_ab.txt
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=(vcs_h264_b200/libvcs_b200.so)
for f in "${libs[@]}"; do
  echo "== $f" >> gpurun_out/dct_ab.txt
  VCS_B200_LIB=$f timeout 300 python tools/prof_dct.py 2>&1 | tail -3 >> gpurun_out/dct_ab.txt
  VCS_B200_LIB=$f timeout 300 python tools/e2e_sched.py "" 2>&1 | tail -1 >> gpurun_out/dct_ab.txt
  VCS_B200_LIB=$f timeout 300 python tools/e2e_sched.py --dense "" 2>&1 | tail -1 >> gpurun_out/dct_ab.txt
  VCS_B200_LIB=$f timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:dct_stage -s 1 -c 2 --csv python tools/prof_dct.py 2>/dev/null | grep dct_stage | awk -F'","' '{print $5, $(NF-2), $(NF)}' >> gpurun_out/dct_ab.txt
done
cat gpurun_out/dct_ab.txt
