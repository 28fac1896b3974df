# round 2, GPU call 4 (2 GPUs): ISP fused kernels (numerics + timing vs NCCL), ISP training parity, MoE fused/grouped tests
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/fused_comm_check.py > gpurun_out/r2_fused_check_n2_v2.log 2>&1; echo "check rc=$?"; grep -E "stage|Error|error" gpurun_out/r2_fused_check_n2_v2.log | tail -30
timeout 600 python -m pytest tests/test_fused_comm_gpu.py tests/test_moe_fused_gpu.py -x -q -k "isp or moe or Moe" > gpurun_out/r2_isp_moe_tests_n2.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_isp_moe_tests_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --tp 2 --tp-mode isp --wp 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_isp_fused1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_isp_fused1.log
B200_ISP_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --tp 2 --tp-mode isp --wp 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_isp_fused0.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_isp_fused0.log
