#!/bin/bash
# round 2, call P: attention Q-staging change + vectorised StreamPool host logic: tests, bench, stream benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_stream_pool.py tests/test_squeezeformer_stream.py tests/test_efficient_stream.py tests/test_gpu_parity.py "tests/test_gpu_full_size.py::test_32x10s_full_batch_ids_bit_exact" tests/test_gpu_full_size.py::test_config3_squeezeformer_64_live_streams -x -q -m gpu > gpurun_out/r02_p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_p_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_p_bench.json 2> gpurun_out/r02_p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_p_bench.json")); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["roofline"]["frac"], {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items()})
PY
for m in squeezeformer conformer efficient_conformer; do timeout 600 python tools/stream_bench.py --model $m --streams 64 2>gpurun_out/r02_p_stream_$m.err | tee gpurun_out/r02_p_stream_$m.json | cut -c1-420; tail -2 gpurun_out/r02_p_stream_$m.err; done
