#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_breakdown.py 2>&1 | grep -v Warning | tee gpurun_out/e2e_breakdown.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 6 -f -o gpurun_out/gemm_prof python tools/gemm_prof.py > gpurun_out/gemm_prof.log 2>&1
tail -3 gpurun_out/gemm_prof.log
ls -la gpurun_out/
