#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -30 | cut -c1-300 | tee gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
