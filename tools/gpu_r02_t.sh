#!/bin/bash
# round 2, call T: final validation — whole -m gpu suite, smoke, bench N=1 (driver defaults)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_t_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_t_smoke.log
timeout 600 python bench.py > gpurun_out/r02_t_bench.json 2> gpurun_out/r02_t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_t_bench.json")); print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["roofline"]["frac"], d["gpu_launches"], d["cpu_baseline"]["value"], d["cpu_baseline"]["token_ids_match_gpu"], {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items()})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
