# round 2, GPU call 7 (2 GPUs): AG / weight-gather with arrival counters (numerics + timing), ISP + TP training parity, ISP bench
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/fused_comm_check.py > gpurun_out/r2_fused_check_n2_v4.log 2>&1; echo "check rc=$?"; grep -E "all_ok|Error" gpurun_out/r2_fused_check_n2_v4.log | tail -3
timeout 900 python -m pytest tests/test_fused_comm_gpu.py -x -q > gpurun_out/r2_fused_tests_n2_v2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_fused_tests_n2_v2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --tp 2 --tp-mode isp --wp 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_isp_v2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_isp_v2.log
