#!/bin/bash
# round 2, final ncu pass of the cta_group::2 build: per-kernel table of one step (launch list incl.) + --set full on the FFN GEMMs
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_aa_step_metrics.csv \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum \
   python tools/profile_step.py > gpurun_out/r02_aa_step_metrics.log 2>&1; echo "ncu metrics rc=$?"
python tools/kernel_roofline.py gpurun_out/r02_aa_step_metrics.csv > gpurun_out/r02_aa_kernel_roofline.md 2>&1; head -12 gpurun_out/r02_aa_kernel_roofline.md
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm_kernel -s 2 -c 2 -f \
   -o gpurun_out/r02_aa_ffn_pair python tools/profile_step.py > gpurun_out/r02_aa_ffn.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02_aa_*
