#!/bin/bash
# round 2, call I: ncu --set full on the prefix beam kernel (config 4) and on the FFN w_1 GEMM; fixed at-size test
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefix_beam -s 2 -c 1 -f -o gpurun_out/r02_i_beam python tools/config_bench.py config4 > gpurun_out/r02_i_beam.log 2>&1; echo "ncu beam rc=$?"
MASR_FUSE_LN=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_kernel -s 2 -c 2 -f \
   -o gpurun_out/r02_i_ffn python tools/profile_step.py > gpurun_out/r02_i_ffn.log 2>&1; echo "ncu ffn rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest "tests/test_gpu_full_size.py::test_32x10s_full_batch_ids_bit_exact" -x -q > gpurun_out/r02_i_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -3 gpurun_out/r02_i_fullsize.log
