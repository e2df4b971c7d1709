#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 python tools/stream_bench.py --model squeezeformer --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_sqz64.json
timeout 600 python tools/stream_bench.py --model conformer --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_conf64.json
timeout 600 python tools/stream_bench.py --model efficient_conformer --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_eff64.json
