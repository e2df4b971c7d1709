#!/bin/bash
# round 2, call M: tcgen05 attention kernel — unit test under a short timeout, then parity + bench A/B
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -x -q -k "relpos_attention" > gpurun_out/r02_m_unit.log 2>&1; echo "unit rc=$?"; tail -25 gpurun_out/r02_m_unit.log | cut -c1-400
if grep -q "passed" gpurun_out/r02_m_unit.log && ! grep -q "failed" gpurun_out/r02_m_unit.log; then
  timeout 600 python -m pytest tests/test_gpu_parity.py "tests/test_gpu_full_size.py::test_32x10s_full_batch_ids_bit_exact" -x -q > gpurun_out/r02_m_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/r02_m_parity.log | cut -c1-300
  for a in tc5 mma; do
    MASR_ATTN=$a timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_m_bench_$a.json 2> gpurun_out/r02_m_bench_$a.err; echo "bench $a rc=$?"
    python - <<PY
import json
d = json.load(open("gpurun_out/r02_m_bench_$a.json")); print("$a", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items() if k in ("attention","qkv_proj","out_proj")})
PY
  done
fi
