#!/bin/bash
# Kernel experiments on the GPU box: bench every library build under greb-climate-model_b200/variants/
# (made with `make EXTRA=-D... OUT=...`, csrc/Makefile) plus the in-tree default, in both arithmetic
# modes.  Usage: tools/variant_bench.sh [members] [steps]  -> gpurun_out/variants.txt
M=${1:-1024}; K=${2:-3}
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for lib in greb-climate-model_b200/libgreb_b200.so greb-climate-model_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  for arith in fast exact; do
    line=$(GREB_B200_LIB=$PWD/$lib timeout 300 python bench.py --members $M --steps $K --warmup 1 --quick --no-e2e --arith $arith 2>>gpurun_out/variants.err | tail -1)
    v=$(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f my/s  %.2f ms/step  sm %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))" 2>/dev/null || echo "FAILED: $line")
    echo "$(basename $lib) $arith $v" | tee -a gpurun_out/variants.txt
  done
done
