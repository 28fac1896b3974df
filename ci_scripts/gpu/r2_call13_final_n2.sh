# round 2, GPU call 13 (2 GPUs): GPU test subsets after the last code change (kernels, training, fused comm / ISP / SP attention, MoE)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_fused_comm_gpu.py tests/test_moe_fused_gpu.py -x -q > gpurun_out/r2_gputests_final_n2.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2_gputests_final_n2.log
