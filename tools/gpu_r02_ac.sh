#!/bin/bash
# round 2: do the pair kernels help or hurt the small-M paths? (stream pool chunk step M = 1024; configs 2/4 shards)
mkdir -p gpurun_out
for m in conformer squeezeformer; do
  for pair in default 0; do
    if [ $pair = default ]; then unset MASR_TC_PAIR; else export MASR_TC_PAIR=$pair; fi
    timeout 300 python tools/stream_bench.py --model $m --streams 64 2>/dev/null | grep "^{" > gpurun_out/r02_ac_stream_${m}_pair${pair}.json
    python -c "import json,sys; d=json.load(open('gpurun_out/r02_ac_stream_${m}_pair${pair}.json')); print('$m pair=$pair', round(d['audio_seconds_per_second']), d['push_latency_ms']['p50'])"
  done
done
unset MASR_TC_PAIR
timeout 600 python tools/config_bench.py config2 config4 config4p config4g config5g squeezeformer 2>/dev/null | grep "^{" > gpurun_out/r02_ac_config_bench.json
python - <<'PY'
import json
for l in open("gpurun_out/r02_ac_config_bench.json"):
    d = json.loads(l); print(d["config"][:60], round(d["audio_seconds_per_second"]), round(d["ms_per_batch"], 2))
PY
