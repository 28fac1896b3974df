#!/usr/bin/env bash
# Weak-scaling sweep of the headline benchmark on one node (both arms).
set -euo pipefail
cd "$(dirname "$0")/.."
for impl in reference ours; do
  for n in 1 2 4 8; do
    if [ "$n" = 1 ]; then python bench.py --impl $impl --gpus 1 --steps 5 --warmup 3
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29578 \
           bench.py --impl $impl --gpus $n --steps 5 --warmup 3; fi
  done
done
