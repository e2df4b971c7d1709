#!/bin/bash
# Dev tool: one gpurun call = kernel tests + parity + GEMM microbench + bench line (logs under gpurun_out/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -15 | tee gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 | tee gpurun_out/t_parity.log
GB_FLAGS=${GB_FLAGS:-5,13} timeout 300 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
