#!/bin/bash
# round 2: cta_group::2 (PAIR) GEMM — bit-identity against the single-CTA kernel, timings, then the GEMM tests and the bench
mkdir -p gpurun_out
timeout 150 python tools/pair_gemm_check.py ffn_w2 > gpurun_out/r02_v_pair_w2.log 2>&1; rc=$?; echo "w2 rc=$rc"; tail -3 gpurun_out/r02_v_pair_w2.log | cut -c1-400
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,memory.used --format=csv; exit 0; fi
timeout 400 python tools/pair_gemm_check.py > gpurun_out/r02_v_pair_all.log 2>&1; rc=$?; echo "all rc=$rc"; cut -c1-330 gpurun_out/r02_v_pair_all.log | tail -20
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py -x -q > gpurun_out/r02_v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_v_tests.log
timeout 300 python bench.py > gpurun_out/r02_v_bench.json 2> gpurun_out/r02_v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_v_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["frac"], d["roofline"].get("launch_us"))
PY
