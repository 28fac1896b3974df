# round 2, GPU call 8 (2 GPUs): reference arm with the tp2 sub-line; 4-layer smokes of the --config paths before the 8-GPU run
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29601 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_ref_n2.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref_n2.log | cut -c1-1800
timeout 200 $T --master-port 29602 bench.py --gpus 2 --config configs/7B_isp_sft.py --tp 2 --wp 2 --seq-len 8192 --micro-bsz 1 --micro-num 2 --segments 2 --layers 4 --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_smoke_isp.log 2>&1; echo "isp rc=$?"; tail -1 gpurun_out/r2_smoke_isp.log | cut -c1-900
timeout 200 $T --master-port 29603 bench.py --gpus 2 --config configs/7B_MoE4_sft.py --layers 4 --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_smoke_moe.log 2>&1; echo "moe rc=$?"; tail -1 gpurun_out/r2_smoke_moe.log | cut -c1-900
timeout 200 $T --master-port 29604 bench.py --gpus 2 --config configs/20B_internlm2.py --tp 1 --pp 2 --layers 4 --micro-num 4 --steps 2 --warmup 3 --no-tp2 > gpurun_out/r2_smoke_20b.log 2>&1; echo "20b rc=$?"; tail -1 gpurun_out/r2_smoke_20b.log | cut -c1-900
