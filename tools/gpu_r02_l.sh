#!/bin/bash
# round 2, call L: pipeline depth 3 — parity tests of the pipelined API, N=1 bench (new roofline timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_evaluate.py "tests/test_gpu_full_size.py::test_32x10s_full_batch_ids_bit_exact" tests/test_vad.py -x -q -m gpu > gpurun_out/r02_l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_l_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_l_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_l_bench.json")); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d["roofline"])
PY
