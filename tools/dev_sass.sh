#!/bin/bash
# Development helper: compile the hot kernels of ONE window radius subset and dump their SASS.
#   tools/dev_sass.sh [sub]    sub = 0:{R0..3} 1:{R4} 2:{R5} 3:{R6,7}   -> /tmp/dev/sass_ssd.txt, /tmp/dev/sass_ncc.txt
SUB=${1:-2}
set -e
mkdir -p /tmp/dev
SRC="$(dirname "$0")/../introtocomputervision_b200/csrc/fast_inst.cu"
for c in 0 1; do
  part=$((c * 4 + SUB))
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DSB_PART=$part -diag-suppress 177,128 -Xptxas -v \
    -cubin -o /tmp/dev/part$c.cubin "$SRC" 2> /tmp/dev/ptxas$c.txt &
done
wait
grep -E "Compiling|registers|spill" /tmp/dev/ptxas0.txt /tmp/dev/ptxas1.txt | sed -e 's/ptxas info    : //' | cut -c1-160
cuobjdump -sass /tmp/dev/part0.cubin | awk '/Function : /{f=$3} { if (f ~ "ELi24ELi8E") print }' > /tmp/dev/sass_ssd.txt
cuobjdump -sass /tmp/dev/part1.cubin | awk '/Function : /{f=$3} { if (f ~ "ELi24ELi8E") print }' > /tmp/dev/sass_ncc.txt
