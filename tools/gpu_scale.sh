#!/bin/bash
# usage: tools/gpu_scale.sh TAG N -> bench.py + the host-link probe on N GPUs of the box (run under gpurun --gpus N)
TAG=$1; N=$2
if [ "$N" = "1" ]; then
  python tools/pcie_probe.py > gpurun_out/pcie_${TAG}_n$N.json 2> gpurun_out/pcie_${TAG}_n$N.err
  python bench.py --gpus 1 --no-cpu-baseline --text-sample 0 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py > gpurun_out/pcie_${TAG}_n$N.json 2> gpurun_out/pcie_${TAG}_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline --text-sample 0 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
fi
cat gpurun_out/pcie_${TAG}_n$N.json; python tools/show_bench.py gpurun_out/scale_${TAG}_n$N.json; nproc; free -g | head -2
