#!/bin/bash
# round 2, call U (2 GPUs): e2e at N=2 under variations — CPU affinity off, eager (not graph-captured) all-gather
mkdir -p gpurun_out
export MASR_BENCH_WATCHDOG_S=240
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_u_$1.json 2> gpurun_out/r02_u_$1.err
  python - <<PY
import json
for l in open("gpurun_out/r02_u_$1.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$1", "value ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "sync-call ms", round(d["e2e"]["sync_call"]["ms_per_step"],3), d["config"].get("collective","")[:40], d["config"].get("cpu_affinity"))
PY
}
run base 29541
MASR_BENCH_AFFINITY=0 run noaff 29542
MASR_GRAPH_GATHER=0 run eagergather 29543
