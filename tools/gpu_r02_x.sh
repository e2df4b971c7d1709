#!/bin/bash
# round 2: kernel forms of tc_gemm (pair / alternating epilogue groups) — bit identity + timings, GEMM tests, bench
mkdir -p gpurun_out
timeout 200 python tools/pair_gemm_check.py ffn_w1 qkv > gpurun_out/r02_x_first.log 2>&1; rc=$?; echo "first rc=$rc"; cut -c1-420 gpurun_out/r02_x_first.log | tail -4
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 python tools/pair_gemm_check.py > gpurun_out/r02_x_all.log 2>&1; rc=$?; echo "all rc=$rc"; cut -c1-420 gpurun_out/r02_x_all.log | tail -20
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py -x -q > gpurun_out/r02_x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_x_tests.log
timeout 300 python bench.py > gpurun_out/r02_x_bench.json 2> gpurun_out/r02_x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_x_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["frac"])
PY
