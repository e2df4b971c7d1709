#!/bin/bash
# round 2, call K: pipelined beam search (second stream) test + config 4 throughput line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r02_k_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_k_tests.log | cut -c1-300
timeout 600 python tools/config_bench.py config4 config4p config4g > gpurun_out/r02_k_config_bench.json 2> gpurun_out/r02_k_config_bench.err; echo "config rc=$?"; cut -c1-330 gpurun_out/r02_k_config_bench.json; tail -3 gpurun_out/r02_k_config_bench.err
