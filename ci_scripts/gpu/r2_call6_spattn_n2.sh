# round 2, GPU call 6 (2 GPUs): sequence-parallel attention with peer K/V - numerics vs the single-rank kernel, timing vs Ulysses;
# fused TP / ISP check again with the deeper push loop
set -x
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/sp_attn_check.py > gpurun_out/r2_sp_attn_n2.log 2>&1; echo "sp rc=$?"; grep -E "sp_attn|Error|error|all_ok" gpurun_out/r2_sp_attn_n2.log | tail -12
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/fused_comm_check.py > gpurun_out/r2_fused_check_n2_v3.log 2>&1; echo "check rc=$?"; grep -E "all_ok" gpurun_out/r2_fused_check_n2_v3.log | tail -2
