#!/bin/bash
# usage: tools/gpu_quick.sh TAG -> short PacBio + ONT bench lines (no CPU legs) into gpurun_out/
TAG=$1
python bench.py --no-cpu-baseline --text-sample 0 --steps 10 --parity-reads 1000 --no-dry > gpurun_out/q_$TAG.json 2> gpurun_out/q_$TAG.err; echo "rc=$?" >> gpurun_out/q_$TAG.err
python bench.py --no-cpu-baseline --text-sample 0 --steps 10 --shape ont --parity-reads 1000 --no-dry > gpurun_out/q_${TAG}_ont.json 2> gpurun_out/q_${TAG}_ont.err; echo "rc=$?" >> gpurun_out/q_${TAG}_ont.err
python tools/show_bench.py gpurun_out/q_$TAG.json gpurun_out/q_${TAG}_ont.json; tail -2 gpurun_out/q_$TAG.err
