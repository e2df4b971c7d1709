#!/bin/bash
# round 2, call B: fused LayerNorm / CTC epilogues — unit tests under a short timeout first, then parity + bench A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q -k "residual_layernorm or ctc_head" > gpurun_out/r02_b_unit.log 2>&1; echo "unit rc=$?"
tail -25 gpurun_out/r02_b_unit.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r02_b_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/r02_b_parity.log
MASR_FUSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_bench_fuse1.json 2> gpurun_out/r02_b_bench_fuse1.err; echo "bench1 rc=$?"
cut -c1-300 gpurun_out/r02_b_bench_fuse1.json; python - <<'PY'
import json
for f in ("gpurun_out/r02_b_bench_fuse1.json",):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["kernel_time_shares"], d["roofline"]["launch_ms"])
    except Exception as e: print(f, e)
PY
MASR_FUSE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_bench_fuse0.json 2> gpurun_out/r02_b_bench_fuse0.err; echo "bench0 rc=$?"
cut -c1-200 gpurun_out/r02_b_bench_fuse0.json
