#!/bin/bash
# usage: tools/gpu_artifacts.sh TAG -> the round's measured artefacts into gpurun_out/ (copied to profiles/ afterwards)
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --shape ont --no-cpu-baseline --text-sample 0 > gpurun_out/bench_${TAG}_ont.json 2> gpurun_out/bench_${TAG}_ont.err
python bench.py --seq8 --no-cpu-baseline --text-sample 0 > gpurun_out/bench_${TAG}_seq8.json 2> gpurun_out/bench_${TAG}_seq8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --no-cpu-baseline --text-sample 0 --steps 2 --warmup 3 --resident-only > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_emit -s 3 -c 1 -o gpurun_out/prof_${TAG}_emit -f \
  python bench.py --no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000 > gpurun_out/ncu_full_$TAG.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/smi_$TAG.log
nproc >> gpurun_out/smi_$TAG.log
