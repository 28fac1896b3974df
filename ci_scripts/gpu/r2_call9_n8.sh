# round 2, GPU call 9 (8 GPUs): headline bench with the tp2 sub-line, fused TP / ZeRO / MoE / ISP kernels vs NCCL at 8 ranks,
# the other BASELINE configs (ISP seq 32768 sp8 x wp8, MoE4 ep4 x edp2, 20B TP4 x PP2), sequence-parallel attention at 8 ranks
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $T --master-port 29701 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_bench_n8.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-2500
timeout 240 $T --master-port 29702 tools/fused_comm_check.py > gpurun_out/r2_fused_check_n8.log 2>&1; echo "check rc=$?"; grep -E "all_ok|Error" gpurun_out/r2_fused_check_n8.log | tail -3
timeout 240 $T --master-port 29703 bench.py --gpus 8 --config configs/7B_isp_sft.py --tp 8 --wp 8 --seq-len 32768 --micro-bsz 1 --micro-num 2 --segments 8 --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_bench_n8_isp32k.log 2>&1; echo "isp rc=$?"; tail -1 gpurun_out/r2_bench_n8_isp32k.log | cut -c1-1200
timeout 240 $T --master-port 29704 bench.py --gpus 8 --config configs/7B_MoE4_sft.py --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_bench_n8_moe4.log 2>&1; echo "moe rc=$?"; tail -1 gpurun_out/r2_bench_n8_moe4.log | cut -c1-1200
timeout 240 $T --master-port 29705 bench.py --gpus 8 --config configs/20B_internlm2.py --micro-num 8 --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_bench_n8_20b.log 2>&1; echo "20b rc=$?"; tail -1 gpurun_out/r2_bench_n8_20b.log | cut -c1-1200
timeout 150 $T --master-port 29706 tools/sp_attn_check.py > gpurun_out/r2_sp_attn_n8.log 2>&1; echo "sp rc=$?"; grep -E "all_ok" gpurun_out/r2_sp_attn_n8.log | tail -1 | cut -c1-900
