# round 2, GPU call 2 (2 GPUs): fused TP kernels on the CTA-pair tile - numerics + timing vs NCCL, training parity, tp2 bench
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/fused_comm_check.py > gpurun_out/r2_fused_check_n2.log 2>&1; echo "check rc=$?"; tail -25 gpurun_out/r2_fused_check_n2.log
timeout 600 python -m pytest tests/test_fused_comm_gpu.py -x -q -k "tp2_fused" > gpurun_out/r2_fused_tests_n2.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_fused_tests_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --tp 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2_tp2_fused1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_tp2_fused1.log
B200_TP_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --tp 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2_tp2_fused0.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_tp2_fused0.log
