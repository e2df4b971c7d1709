#!/bin/bash
# round 2: LayerNorm-prologue GEMM (masr_gemm_tc_lnpre_f16x2): bit identity vs separate launches, then the step A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q -k "layernorm_prologue" > gpurun_out/r02_ad_test.log 2>&1; rc=$?; echo "lnpre test rc=$rc"; tail -4 gpurun_out/r02_ad_test.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|error" gpurun_out/r02_ad_test.log | head -10; exit 0; fi
timeout 200 python tools/step_ab.py default= lnpre=MASR_FUSE_LNPRE:1 > gpurun_out/r02_ad_step_ab.json 2> gpurun_out/r02_ad_step_ab.err; echo "ab rc=$?"; cat gpurun_out/r02_ad_step_ab.json; tail -2 gpurun_out/r02_ad_step_ab.err
