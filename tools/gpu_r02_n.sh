#!/bin/bash
# round 2, call N: the whole -m gpu suite on the current code + smoke + sanitizers on the new kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_n_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02_n_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_n_smoke.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_n_memcheck.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc_gemm.py tests/test_beam.py -q -x -m gpu -k "relpos_attention or residual_layernorm or ctc_head or streaming_beam_equals" > gpurun_out/r02_n_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_n_memcheck.log; tail -2 gpurun_out/r02_n_memcheck_stdout.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_n_racecheck.log python -m pytest tests/test_gpu_kernels.py tests/test_beam.py tests/test_deepspeech2.py -q -x -m gpu -k "relpos_attention or streaming_beam_equals or gpu_engine_matches" > gpurun_out/r02_n_racecheck_stdout.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_n_racecheck.log; tail -2 gpurun_out/r02_n_racecheck_stdout.log
