#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | tail -6 | cut -c1-300 | tee gpurun_out/t_all.log
MASR_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_pdl0.json | cut -c1-220
MASR_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-220
tail -3 gpurun_out/bench.err
