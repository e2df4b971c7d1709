#!/bin/bash
# round 2, call D: ncu --set full on the LN-fused w_2 GEMM (4th tc_gemm launch of the step) and its unfused counterpart; new beam tests
mkdir -p gpurun_out
MASR_FUSE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_kernel -s 3 -c 1 -f \
   -o gpurun_out/r02_d_w2_fused python tools/profile_step.py > gpurun_out/r02_d_w2_fused.log 2>&1; echo "ncu fused rc=$?"
MASR_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_kernel -s 3 -c 1 -f \
   -o gpurun_out/r02_d_w2_unfused python tools/profile_step.py > gpurun_out/r02_d_w2_unfused.log 2>&1; echo "ncu unfused rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_beam.py -x -q -m gpu > gpurun_out/r02_d_beam.log 2>&1; echo "beam rc=$?"; tail -15 gpurun_out/r02_d_beam.log
