#!/bin/bash
# round 2, call E: residual prefetch + persistent LSTM + streaming beam: tests, bench, config / stream benches with oracle checks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py tests/test_deepspeech2.py tests/test_beam.py tests/test_evaluate.py -x -q -m gpu > gpurun_out/r02_e_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_e_tests.log
timeout 600 python -m pytest "tests/test_gpu_full_size.py::test_32x10s_full_batch_ids_bit_exact" tests/test_gpu_full_size.py::test_config5_shard_conformer_64_utterances_1_to_30s -x -q > gpurun_out/r02_e_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -4 gpurun_out/r02_e_fullsize.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_e_bench.json 2> gpurun_out/r02_e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_e_bench.json")); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d["roofline"]["frac"], {k: round(v*d["ms_per_step"]*1000) for k, v in d["kernel_time_shares"].items()}, d["cpu_baseline"])
PY
MASR_TC_PRERES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_e_bench_nopreres.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_e_bench_nopreres.json')); print('no preres', round(d['ms_per_step'],3))"
timeout 900 python tools/config_bench.py > gpurun_out/r02_e_config_bench.json 2> gpurun_out/r02_e_config_bench.err; echo "config rc=$?"; cat gpurun_out/r02_e_config_bench.json | cut -c1-260; tail -3 gpurun_out/r02_e_config_bench.err
for m in squeezeformer conformer; do timeout 600 python tools/stream_bench.py --model $m --streams 64 2>gpurun_out/r02_e_stream_$m.err | tee gpurun_out/r02_e_stream_$m.json | cut -c1-400; done
