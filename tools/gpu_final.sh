#!/bin/bash
# usage: tools/gpu_final.sh TAG -> the round's artefacts into gpurun_out/: gpu test log, smoke, both bench arms, the ONT line, launch lists of both shapes
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?" >> gpurun_out/bench_$TAG.err
python bench.py --shape ont --no-cpu-baseline --text-sample 0 --parity-reads 1500 > gpurun_out/bench_${TAG}_ont.json 2> gpurun_out/bench_${TAG}_ont.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_ont.err
bash tools/gpu_launches.sh ${TAG} > gpurun_out/launch_shares_$TAG.txt 2>&1
bash tools/gpu_launches.sh ${TAG}_ont ont > gpurun_out/launch_shares_${TAG}_ont.txt 2>&1
tail -3 gpurun_out/pytest_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; python tools/show_bench.py gpurun_out/bench_$TAG.json gpurun_out/bench_${TAG}_ont.json; cut -c1-300 gpurun_out/bench_${TAG}_reference.json
