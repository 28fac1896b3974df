# round 2, GPU call 12 (8 GPUs): headline bench (dp8, ZeRO overlapped) + the tp2-dp4 sub-line
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-2600
