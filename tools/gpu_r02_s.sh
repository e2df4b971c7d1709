#!/bin/bash
# round 2, call S: fused LayerNorm epilogues in the stream pools (Conformer: pre-norm LN / LN2, Squeezeformer: post-norm + ada)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q -k "postln or residual_layernorm" > gpurun_out/r02_s_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r02_s_unit.log
timeout 900 python -m pytest tests/test_gpu_stream_pool.py tests/test_squeezeformer_stream.py tests/test_efficient_stream.py tests/test_gpu_parity.py tests/test_gpu_full_size.py::test_config3_squeezeformer_64_live_streams tests/test_gpu_predictor_models.py -x -q -m gpu > gpurun_out/r02_s_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_s_tests.log | cut -c1-300
for f in 1 0; do for m in squeezeformer conformer; do MASR_POOL_FUSE_LN=$f timeout 600 python tools/stream_bench.py --model $m --streams 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fuse=$f', '$m', round(d['audio_seconds_per_second']), d['push_latency_ms'], d['kernel_launches_per_push'], d['streams_equal_single_stream_path'])"; done; done
