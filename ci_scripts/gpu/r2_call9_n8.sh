# round 2, GPU call 9 (8 GPUs, hard cap 8 min): headline bench with the tp2 sub-line, fused TP / ZeRO / MoE / ISP kernels vs NCCL
# at 8 ranks, the other BASELINE configs (ISP seq 32768 sp8 x wp8, MoE4 ep4 x edp2, 20B TP4 x PP2), SP attention at 8 ranks
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() {  # name, timeout, command...: every step has its own limit and the whole script stops at the 8-minute mark
  local name=$1 lim=$2; shift 2
  local left=$(( 470 - SECONDS )); [ $left -lt 20 ] && { echo "$name: skipped (time)"; return; }
  [ $lim -gt $left ] && lim=$left
  timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? t=${SECONDS}s"; tail -1 gpurun_out/$name.log | cut -c1-2200
}
run r2_bench_n8 260 $T --master-port 29701 bench.py --gpus 8 --steps 4 --warmup 3
run r2_fused_check_n8 120 $T --master-port 29702 tools/fused_comm_check.py
run r2_bench_n8_isp32k 110 $T --master-port 29703 bench.py --gpus 8 --config configs/7B_isp_sft.py --tp 8 --wp 8 --seq-len 32768 --micro-bsz 1 --micro-num 2 --segments 8 --steps 2 --warmup 3 --no-tp2
run r2_bench_n8_moe4 110 $T --master-port 29704 bench.py --gpus 8 --config configs/7B_MoE4_sft.py --steps 2 --warmup 3 --no-tp2
run r2_bench_n8_20b 110 $T --master-port 29705 bench.py --gpus 8 --config configs/20B_internlm2.py --micro-num 8 --steps 2 --warmup 3 --no-tp2
run r2_sp_attn_n8 80 $T --master-port 29706 tools/sp_attn_check.py
