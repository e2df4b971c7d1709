#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 900 python tools/config_bench.py 2>&1 | grep -v Warn | tee gpurun_out/config_bench.json
