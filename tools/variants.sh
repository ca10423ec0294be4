#!/bin/bash
# usage (here): tools/variants.sh build NAME "-DX=1 -DY=2" ...   -> build/variants/libtc_NAME.so (tuning builds of the CUDA library)
#       (GPU):  tools/variants.sh run NAME...                     -> short PacBio + ONT bench lines per variant into gpurun_out/var_NAME*.json
cmd=$1; shift
if [ "$cmd" = build ]; then
  mkdir -p build/variants
  while [ $# -ge 2 ]; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -Xcompiler -pthread $2 \
      -o build/variants/libtc_$1.so transcriptclean_b200/csrc/tc_b200.cu transcriptclean_b200/csrc/tc_host.cpp transcriptclean_b200/csrc/tc_tables.cpp -Xptxas -v 2>&1 | grep -A1 "${VKERNEL:-k_verify}" | grep spill
    shift 2
  done
else
  for v in "$@"; do
    TC_B200_LIB=$PWD/build/variants/libtc_$v.so python bench.py --no-cpu-baseline --text-sample 0 --steps 10 --parity-reads 500 --no-dry > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
    TC_B200_LIB=$PWD/build/variants/libtc_$v.so python bench.py --no-cpu-baseline --text-sample 0 --steps 10 --parity-reads 500 --no-dry --shape ont > gpurun_out/var_${v}_ont.json 2>gpurun_out/var_${v}_ont.err
    python tools/show_bench.py gpurun_out/var_$v.json gpurun_out/var_${v}_ont.json
  done
fi
