#!/bin/bash
# round 2, call G: persistent LSTM v2 (cp.async ring + FFMA2), exact beam tests, stream benches, DS2 / beam config lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deepspeech2.py tests/test_beam.py tests/test_gpu_configs.py tests/test_vad.py tests/test_gpu_full_size.py::test_config5_shard_conformer_64_utterances_1_to_30s tests/test_gpu_full_size.py::test_config4_shard_efficient_conformer_32x10s_nonstreaming -x -q -m gpu > gpurun_out/r02_g_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02_g_tests.log
timeout 600 python tools/config_bench.py deepspeech2 config4 config5 > gpurun_out/r02_g_config_bench.json 2> gpurun_out/r02_g_config_bench.err; echo "config rc=$?"; cut -c1-230 gpurun_out/r02_g_config_bench.json; tail -3 gpurun_out/r02_g_config_bench.err
MASR_LSTM_PERSISTENT=0 timeout 600 python tools/config_bench.py deepspeech2 2>/dev/null | cut -c1-200
for m in squeezeformer conformer efficient_conformer; do timeout 600 python tools/stream_bench.py --model $m --streams 64 2>gpurun_out/r02_g_stream_$m.err | tee gpurun_out/r02_g_stream_$m.json | cut -c1-500; tail -2 gpurun_out/r02_g_stream_$m.err; done
