#!/bin/bash
# usage: tools/gpu_text.sh TAG -> SAM-text path throughput over worker counts / block sizes
TAG=$1
for CFG in "1 400" "3 16" "3 48" "3 100" "2 100" "4 48"; do
  set -- $CFG
  TC_HOST_TIMING=$3 python bench.py --no-cpu-baseline --steps 3 --parity-reads 0 --no-dry --reads 300000 --text-sample 250000 --text-workers $1 --text-block-mb $2 > gpurun_out/text_${TAG}_w$1_b$2.json 2> gpurun_out/text_${TAG}_w$1_b$2.err
  python -c "import json;d=json.loads(open('gpurun_out/text_${TAG}_w$1_b$2.json').read().strip().splitlines()[-1]);print('workers $1 block $2 MB:', round(d['sam_text_path']['value']), d['sam_text_path']['sample'][60:100])"
done
