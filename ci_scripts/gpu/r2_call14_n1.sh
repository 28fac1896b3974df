# round 2, GPU call 14 (1 GPU): what the driver runs at round end - pytest -m gpu, smoke(), bench.py N = 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_n1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gputests_n1.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n1.log | cut -c1-1800
