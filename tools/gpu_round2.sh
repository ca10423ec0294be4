#!/bin/bash
# usage: tools/gpu_round2.sh TAG [skiptests] -> gpu tests, both bench arms, PacBio + ONT lines into gpurun_out/
TAG=$1
if [ "$2" != "skiptests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
fi
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?" >> gpurun_out/bench_$TAG.err
python bench.py --no-cpu-baseline --text-sample 0 --shape ont --parity-reads 1500 > gpurun_out/bench_${TAG}_ont.json 2> gpurun_out/bench_${TAG}_ont.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_ont.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
tail -3 gpurun_out/pytest_$TAG.log; cut -c1-600 gpurun_out/bench_$TAG.json
