# round 2, GPU call 3 (1 GPU): grouped GEMM numerics, MoE layers on the grouped path, kernel regression tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q > gpurun_out/r2_kernels_n1.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_kernels_n1.log
timeout 300 python tools/moe_bench.py --help > /dev/null 2>&1; echo "moe_bench help rc=$?"
