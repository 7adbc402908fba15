#!/bin/bash
# Warp-placement experiments on the GPU box (greb_b200.cu greb_layouts, env GREB_B200_LAYOUT):
# bench every placement in both arithmetic modes.  Usage: tools/layout_bench.sh [members] [steps]
M=${1:-1024}; K=${2:-3}
mkdir -p gpurun_out
: > gpurun_out/layouts.txt
L0="0,1,2,3,4,5,6,7,8,9,10,11,12,13,15,15"; O0="2,3,6,7,10,11,0,1,4,5,8,9"
L1="0,1,2,12,3,4,5,13,6,7,8,15,9,10,11,15"; O1="0,3,1,4,2,5,6,9,7,10,8,11"
L2="0,1,2,12,3,4,5,13,6,7,8,9,10,15,15,11"; O2="1,4,2,5,9,7,8,11,0,3,6,10"
for lay in 0 1 2 3 "$L0/0/$O0" "$L1/0/$O1" "$L2/0/$O2" "$L2/fff/$O2" "$L0/f0f/$O0"; do
  for arith in exact fast; do
    line=$(GREB_B200_LAYOUT="$lay" timeout 300 python bench.py --members $M --steps $K --warmup 1 --quick --no-e2e --arith $arith 2>>gpurun_out/layouts.err | tail -1)
    v=$(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f my/s  %.2f ms/step' % (d['value'], d['ms_per_step']))" 2>/dev/null || echo "FAILED: $line")
    echo "layout $lay $arith $v" | tee -a gpurun_out/layouts.txt
  done
done
