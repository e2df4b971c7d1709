#!/bin/bash
# Dev tool: everything one round measures on a B200, in one `gpurun` call (outputs under gpurun_out/; copy what should be
# judged into profiles/).  Usage: gpurun --timeout 2700 -- 'bash tools/gpu_round.sh [tests|bench|ncu|stream|configs ...]'
set -u
mkdir -p gpurun_out
what=${*:-tests bench ncu stream configs}
for w in $what; do
  case $w in
    tests)   timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -8 | cut -c1-300 | tee gpurun_out/t_all.log ;;
    bench)   timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    ncu)     # per-launch metrics of one step (-> tools/kernel_roofline.py) and --set full captures of the SIMT kernels + FFN GEMMs
             timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics.csv \
               --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum \
               python tools/profile_step.py > gpurun_out/step_metrics.log 2>&1
             timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
               -k regex:"fbank_kernel|dwconv_ln_silu|relpos_attention_mma|conv1_cmvn|layernorm2|ctc_frame_argmax|ctc_greedy_collapse" -c 9 -f \
               -o gpurun_out/step_kernels python tools/profile_step.py > gpurun_out/step_kernels.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 4 -f -o gpurun_out/gemm_prof \
               env GP_FLAGS=5 python tools/gemm_prof.py > gpurun_out/gemm_prof.log 2>&1 ;;
    stream)  for m in squeezeformer conformer efficient_conformer; do
               timeout 600 python tools/stream_bench.py --model $m --streams 64 2>&1 | grep -v Warn | tee gpurun_out/stream_$m.json; done ;;
    configs) timeout 900 python tools/config_bench.py 2>&1 | grep -v Warn | tee gpurun_out/config_bench.json ;;
  esac
done
