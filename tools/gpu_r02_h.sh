#!/bin/bash
# round 2, call H: the whole -m gpu suite, bench, beam configs after the beam-kernel rework, per-kernel ncu table of one step
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_h_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02_h_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_h_bench.json 2> gpurun_out/r02_h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_h_bench.json")); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d["roofline"]["frac"], {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items()})
PY
timeout 600 python tools/config_bench.py config4 config4g config5 config5g > gpurun_out/r02_h_config_bench.json 2> gpurun_out/r02_h_config_bench.err; echo "config rc=$?"; cut -c1-230 gpurun_out/r02_h_config_bench.json; tail -3 gpurun_out/r02_h_config_bench.err
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_h_step_metrics.csv \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum \
   python tools/profile_step.py > gpurun_out/r02_h_step_metrics.log 2>&1; echo "ncu rc=$?"
python tools/kernel_roofline.py gpurun_out/r02_h_step_metrics.csv > gpurun_out/r02_h_kernel_roofline.md 2>&1; head -20 gpurun_out/r02_h_kernel_roofline.md
for m in squeezeformer; do timeout 600 python tools/stream_bench.py --model $m --streams 64 2>gpurun_out/r02_h_stream_$m.err | tee gpurun_out/r02_h_stream_$m.json | cut -c1-600; tail -2 gpurun_out/r02_h_stream_$m.err; done
