#!/bin/bash
# usage: tools/gpu_prof_multi.sh TAG REGEX COUNT [shape] -> ncu --set full of the first COUNT launches matching REGEX
TAG=$1; K=$2; C=$3; SHAPE=${4:-pacbio}
ncu --set full --clock-control none --import-source on -k regex:$K -c $C -o gpurun_out/prof_$TAG -f \
  python bench.py --no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000 --shape $SHAPE > gpurun_out/ncu_$TAG.log 2>&1
