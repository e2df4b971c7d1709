#!/bin/bash
# round 2: compute-sanitizer memcheck + racecheck on the cta_group::2 kernels and the LayerNorm-prologue GEMM (small shapes)
mkdir -p gpurun_out
SEL='(pair_kernel and (385 or 129)) or (layernorm_prologue and (1000-768 or 129-2048 or 77-512))'
timeout 100 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_ae_memcheck.log python -m pytest tests/test_gpu_tc_gemm.py -q -x -m gpu -k "$SEL" > gpurun_out/r02_ae_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_ae_memcheck.log; tail -1 gpurun_out/r02_ae_memcheck_stdout.log
timeout 100 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_ae_racecheck.log python -m pytest tests/test_gpu_tc_gemm.py -q -x -m gpu -k "$SEL" > gpurun_out/r02_ae_racecheck_stdout.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_ae_racecheck.log; tail -1 gpurun_out/r02_ae_racecheck_stdout.log
