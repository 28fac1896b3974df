# 1 GPU, last call of round 2: per-kernel roofline with programmatic dependent launch OFF (kernel durations without the
# dependency wait), then the driver's GPU tests and smoke() on the final tree
mkdir -p gpurun_out
B200_PDL=0 timeout 110 python tools/roofline.py --out gpurun_out/roofline_r2_nopdl > gpurun_out/r2_roofline_nopdl.log 2>&1; echo "roofline rc=$?"; tail -22 gpurun_out/r2_roofline_nopdl.log
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_n1_v3.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gputests_n1_v3.log
timeout 45 python __graft_entry__.py smoke 2>&1 | tail -2
