#!/bin/bash
# Build A/B variants of the library with different wrap8 subtract splits (ALU : FMA) into build/variants/.
# Each argument is "MASK DEN": offsets d with bit (d mod DEN) set in MASK subtract on the ALU pipe.
# Usage: tools/ab_me.sh "0x3 5" "0x1 2" "0x1 3" ...   then on the GPU box:
#   for f in build/variants/*.so; do VCS_B200_LIB=$f python tools/prof_me.py; done
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  set -- $v
  out=build/variants/libvcs_alu$1of$2.so
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    -Xcompiler -fPIC -shared -cudart static -I include -DVCS_WRAP_ALU_MASK=$1 -DVCS_WRAP_ALU_DEN=$2 \
    -o $out vcs_h264_b200/csrc/vcs_b200.cu &
done
wait
ls -la build/variants/
