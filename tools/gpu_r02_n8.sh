#!/bin/bash
# round 2: the bench at N GPUs (torchrun), as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
export MASR_BENCH_WATCHDOG_S=240
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "n$N rc=$?"
tail -5 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "verified", d.get("gather_verified"), "strong", d.get("strong_scaling"), d["config"].get("collective"), d["config"].get("cpu_affinity"))
PY
