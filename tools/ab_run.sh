#!/bin/bash
# GPU-side A/B of search-kernel builds (tools/ab_me.sh or any -D variant under build/variants/): GPU tests on the in-tree
# library first, then tools/prof_me.py (45 P-frames, both costs) twice per library.  Output: gpurun_out/ab_me.txt
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/ab_me.txt
for rep in 1 2; do
for f in vcs_h264_b200/libvcs_b200.so build/variants/*.so; do
  echo "== $f" >> gpurun_out/ab_me.txt
  VCS_B200_LIB=$f T=46 timeout 120 python tools/prof_me.py 2>&1 | grep "metric" >> gpurun_out/ab_me.txt
done
done
cat gpurun_out/ab_me.txt
