#!/bin/bash
# Dev tool: one gpurun call = new tests + bench + per-kernel ncu metrics of one step (logs under gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_squeezeformer_stream.py tests/test_gpu_parity.py tests/test_gpu_kernels.py tests/test_gpu_predictor_models.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/t_new.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum python tools/profile_step.py > gpurun_out/step_metrics.log 2>&1
tail -2 gpurun_out/step_metrics.log
