#!/bin/bash
# round 2, call J: beam kernel rework v2 (parallel digit scan, shuffle bitonic): exact tests + config timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_beam.py tests/test_gpu_configs.py tests/test_gpu_full_size.py::test_config5_shard_conformer_64_utterances_1_to_30s tests/test_gpu_full_size.py::test_config4_shard_efficient_conformer_32x10s_nonstreaming -x -q -m gpu > gpurun_out/r02_j_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_j_tests.log | cut -c1-300
timeout 600 python tools/config_bench.py config4 config4g config5 config5g > gpurun_out/r02_j_config_bench.json 2> gpurun_out/r02_j_config_bench.err; echo "config rc=$?"; cut -c1-230 gpurun_out/r02_j_config_bench.json; tail -3 gpurun_out/r02_j_config_bench.err
