# 1 GPU: per-kernel roofline of one step (torch.profiler), then the driver's round-end sequence (GPU tests, smoke, short bench)
mkdir -p gpurun_out
timeout 170 python tools/roofline.py --out gpurun_out/roofline_r2 > gpurun_out/r2_roofline.log 2>&1; echo "roofline rc=$?"; tail -22 gpurun_out/r2_roofline.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_n1_v2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gputests_n1_v2.log
timeout 90 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 120 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_v2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n1_v2.log
