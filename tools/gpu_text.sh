#!/bin/bash
# usage: tools/gpu_text.sh TAG "W B" ... -> SAM-text path (250k reads) for the given worker counts / block sizes, with per-stage worker times
TAG=$1; shift
for CFG in "$@"; do
  set -- $CFG
  TC_STREAM_TIMING=1 python bench.py --no-cpu-baseline --steps 3 --parity-reads 0 --no-dry --reads 300000 --text-sample 250000 --text-workers $1 --text-block-mb $2 > gpurun_out/text_${TAG}_w$1_b$2.json 2> gpurun_out/text_${TAG}_w$1_b$2.err
  python -c "import json;d=json.loads(open('gpurun_out/text_${TAG}_w$1_b$2.json').read().strip().splitlines()[-1]);print('workers $1 block $2 MB:', round(d['sam_text_path']['value']))"
  grep "correct_sam_stream worker" gpurun_out/text_${TAG}_w$1_b$2.err | tail -$1
done
