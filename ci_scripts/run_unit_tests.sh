#!/usr/bin/env bash
# CPU tier: every test that does not need a GPU (gloo, multi-process).  GPU tier: kernel numerics + training smoke.
set -euo pipefail
cd "$(dirname "$0")/.."
python internevo_b200/csrc/build.py
python -m pytest tests -x -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
    python -m pytest tests -x -q -m gpu
fi
