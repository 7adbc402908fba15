#!/bin/bash
# strong-scaling sweep of the 0.25-degree big-grid path on one 8-GPU box (profiles/r01_config5_*.json)
mkdir -p gpurun_out
run() {  # n_gpus period
  if [ "$1" = 1 ]; then
    timeout 200 python tools/run_bigrid.py --steps 0.25 --period $2 > gpurun_out/bigrid_n$1_s$2.log 2>&1
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 \
      tools/run_bigrid.py --steps 0.25 --period $2 > gpurun_out/bigrid_n$1_s$2.log 2>&1
  fi
  echo "n=$1 s=$2 rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/bigrid_n$1_s$2.log | head -1
}
run 1 16; run 2 16; run 4 16; run 8 16; run 8 32
