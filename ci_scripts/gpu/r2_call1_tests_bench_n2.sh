set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_n2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_gputests_n2.log; tail -5 gpurun_out/r2_gputests_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2_overlap.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_bench_n2_overlap.log
B200_ZERO_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-tp2 > gpurun_out/r2_bench_n2_serial.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2_serial.log
