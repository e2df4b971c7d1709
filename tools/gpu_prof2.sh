#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fbank_kernel|dwconv_ln_silu|relpos_attention_mma|conv1_cmvn|layernorm2|ctc_frame_argmax" -c 8 -f -o gpurun_out/step_kernels python tools/profile_step.py > gpurun_out/step_kernels.log 2>&1
tail -3 gpurun_out/step_kernels.log
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum python tools/profile_step.py > gpurun_out/step_metrics.log 2>&1
tail -2 gpurun_out/step_metrics.log
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t_configs.log
ls -la gpurun_out/
