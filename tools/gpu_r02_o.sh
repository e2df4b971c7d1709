#!/bin/bash
# round 2, call O: dwconv TW A/B, stream push breakdown, ncu capture of the tcgen05 attention kernel
mkdir -p gpurun_out
for tw in 4 8; do
  MASR_DW_TW=$tw timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k dwconv > /dev/null 2>&1; echo "dwconv tw=$tw tests rc=$?"
  MASR_DW_TW=$tw timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_o_bench_tw$tw.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_o_bench_tw$tw.json")); print("tw=$tw", round(d["ms_per_step"],3), {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items() if k in ("dwconv_ln_silu","attention")})
PY
done
timeout 300 python tools/stream_breakdown.py 2>/dev/null | tee gpurun_out/r02_o_stream_breakdown.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:relpos_attention_tc5 -c 1 -f -o gpurun_out/r02_o_attn python tools/profile_step.py > gpurun_out/r02_o_attn.log 2>&1; echo "ncu attn rc=$?"
