# round 2, GPU call 10 (1 GPU): ncu --set full captures of attention fwd / bwd and the grouped GEMM (one launch each)
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_fwd_kernel -s 2 -c 1 -f -o gpurun_out/attn_fwd_r2 python tools/run_attn_once.py > gpurun_out/ncu_attn_fwd.log 2>&1; echo "rc=$?"
timeout 300 $NCU -k regex:attn_bwd_kernel -s 2 -c 1 -f -o gpurun_out/attn_bwd_r2 python tools/run_attn_once.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "rc=$?"
timeout 300 $NCU -k regex:gemm_bf16_kernel -s 6 -c 1 -f -o gpurun_out/grouped_gemm_r2 python tools/run_grouped_once.py > gpurun_out/ncu_grouped.log 2>&1; echo "rc=$?"
python tools/run_grouped_once.py 2>&1 | tail -1
timeout 120 python tools/kernel_check.py attn 2>&1 | tail -6
ls -la gpurun_out/*.ncu-rep
timeout 200 python tools/kernel_check.py gemm 2>&1 | tail -4
timeout 400 python tools/kernel_check.py gemm_perf 2>&1 | tail -12
