#!/bin/bash
# N-rank probe of the fused distributed H.v: tiled TMA path vs gather path (QR_P2P_TILE=0), via tools/multi_gpu_config.py
# usage: tools/p2p_probe.sh <nproc> <cfg> [tag]
N=$1; CFG=$2; TAG=${3:-p2p}
mkdir -p gpurun_out
for tile in 1 0; do
  QR_P2P_TILE=$tile timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      tools/multi_gpu_config.py $CFG 2> gpurun_out/${TAG}_${CFG}_n${N}_tile${tile}.err | grep '^{' | tee -a gpurun_out/${TAG}_${CFG}_n${N}.jsonl
  tail -3 gpurun_out/${TAG}_${CFG}_n${N}_tile${tile}.err
done
