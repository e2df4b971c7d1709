#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_pool.py tests/test_squeezeformer_stream.py tests/test_efficient_stream.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/t_stream.log
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k config5 2>&1 | grep -v Warning | tail -40 | cut -c1-400 | tee gpurun_out/t_configs.log
timeout 600 python tools/stream_bench.py --model squeezeformer --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_sqz64.json
timeout 600 python tools/stream_bench.py --model conformer --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_conf64.json
