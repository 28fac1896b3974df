# round 2, GPU call 11 (4 GPUs): headline bench (dp4, ZeRO overlapped) + the tp2-dp2 sub-line after a fresh-store re-initialisation
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r2_bench_n4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n4.log | cut -c1-2600
