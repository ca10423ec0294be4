#!/bin/bash
# usage: tools/gpu_check.sh TAG  -> gpu parity tests + PacBio/ONT benches into gpurun_out/
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
python bench.py --no-cpu-baseline --text-sample 0 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --no-cpu-baseline --text-sample 0 --shape ont > gpurun_out/bench_${TAG}_ont.json 2> gpurun_out/bench_${TAG}_ont.err
