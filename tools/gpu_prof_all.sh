#!/bin/bash
# usage: tools/gpu_prof_all.sh TAG [shape] [kernel-regex] -> ncu --set full of one launch of each kernel of a device-resident step (120k reads)
TAG=$1; SHAPE=${2:-pacbio}; K=${3:-'^k_(measure|plan|jn_init|jn_fix|small|verify|emit)'}
ARGS="--no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000 --shape $SHAPE --resident-only --parity-reads 0"
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || exit 1
N=$(python -c "import re;print(len(re.findall(r'\|', '$K'))+1)")
ncu --set full --clock-control none --import-source on -k "regex:$K" -s $((N*3)) -c $N -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_$TAG.log 2>&1
