#!/bin/bash
# round 2, call F (2 GPUs): bench at N=2 — all-gather captured in the step graph, gather_verified, strong-scaling leg
mkdir -p gpurun_out
export MASR_BENCH_WATCHDOG_S=300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_f_bench_n2.json 2> gpurun_out/r02_f_bench_n2.err; echo "n2 rc=$?"
cut -c1-1500 gpurun_out/r02_f_bench_n2.json; tail -15 gpurun_out/r02_f_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_f_bench_n2.json")); print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "verified", d.get("gather_verified"), "strong", d.get("strong_scaling"), d["config"].get("collective"))
except Exception as e: print("parse failed", e)
PY
