#!/bin/bash
# usage: tools/gpu_prof.sh TAG [kernel-regex] [shape] -> one ncu --set full capture of the kernel (120k reads)
TAG=$1; K=${2:-k_emit}; SHAPE=${3:-pacbio}
python bench.py --no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000 --shape $SHAPE > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_$TAG -f \
  python bench.py --no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000 --shape $SHAPE > gpurun_out/ncu_$TAG.log 2>&1
