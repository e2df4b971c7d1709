#!/bin/bash
# round 2, call C: optimized fused epilogues — unit tests, bench A/B, ncu launch lists (fused / unfused), full capture of the LN-fused GEMMs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q -k "residual_layernorm or ctc_head" > gpurun_out/r02_c_unit.log 2>&1; echo "unit rc=$?"
tail -5 gpurun_out/r02_c_unit.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_c_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r02_c_parity.log
for f in 1 0; do
  MASR_FUSE=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c_bench_fuse$f.json 2> gpurun_out/r02_c_bench_fuse$f.err; echo "bench fuse=$f rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_c_bench_fuse$f.json")); print("fuse=$f", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items()})
PY
  MASR_FUSE=$f timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_c_launches_fuse$f.csv \
     --metrics gpu__time_duration.sum python tools/profile_step.py > gpurun_out/r02_c_launches_fuse$f.log 2>&1
  python tools/summarize_launches.py gpurun_out/r02_c_launches_fuse$f.csv | head -24
done
MASR_FUSE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_gemm_kernel<false, 16, true>|tc_gemm_kernel<0, 16, 1>" -c 4 -f \
   -o gpurun_out/r02_c_lngemm python tools/profile_step.py > gpurun_out/r02_c_lngemm.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
