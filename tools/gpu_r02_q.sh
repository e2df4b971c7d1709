#!/bin/bash
# round 2, call Q: final per-kernel ncu table of one step (tc5 attention build) + launch list
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_q_step_metrics.csv \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum \
   python tools/profile_step.py > gpurun_out/r02_q_step_metrics.log 2>&1; echo "ncu rc=$?"
python tools/kernel_roofline.py gpurun_out/r02_q_step_metrics.csv > gpurun_out/r02_q_kernel_roofline.md 2>&1; head -22 gpurun_out/r02_q_kernel_roofline.md
python tools/summarize_launches.py gpurun_out/r02_q_step_metrics.csv > gpurun_out/r02_q_launch_summary.md 2>&1; tail -3 gpurun_out/r02_q_launch_summary.md
