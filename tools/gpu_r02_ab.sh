#!/bin/bash
# round 2 final validation of the cta_group::2 build: all -m gpu tests, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_ab_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_ab_smoke.log
timeout 400 python bench.py > gpurun_out/r02_ab_bench.json 2> gpurun_out/r02_ab_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_ab_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
PY
