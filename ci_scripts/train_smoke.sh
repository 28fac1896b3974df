#!/usr/bin/env bash
# End-to-end training smoke: demo config, 2 ranks, loss must fall (exit code of train.py) and a checkpoint must appear.
set -euo pipefail
cd "$(dirname "$0")/.."
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29577 \
    train.py --config configs/demo.py --launcher torch
