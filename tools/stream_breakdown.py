"""Dev tool: where the time of one `StreamPool.push` (64 streams x 0.5 s) goes — host phases timed with a device synchronise
after each (so the sum exceeds the un-instrumented push), plus the pure GPU time of the chunk step's CUDA graph."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from masr_b200 import synth
from masr_b200.squeezeformer import SqueezeformerEngine
from masr_b200.stream_pool import StreamPool

S, PUSH, N = 64, 8000, 30
eng = SqueezeformerEngine(synth.squeezeformer_state_dict(0, streaming=True), streaming=True)
total = (N + 6) * PUSH
pcm = [(np.clip(synth.noise_audio(500 + s, total), -1, 1) * 32767).astype("<i2") for s in range(S)]
pool = StreamPool(eng, synth.vocabulary(), n_slots=S, max_frames=((total // 160) // 4 + 64))
T = {}


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


eng.fbank = timed("fbank (+H2D of 64 x 8000 samples)", eng.fbank)
pool.pool.step = timed("pool.step (meta H2D + CUDA graph of the chunk step)", pool.pool.step)
for k in range(6):
    pool.push({s: pcm[s][k * PUSH:(k + 1) * PUSH].tobytes() for s in range(S)})
T.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(6, 6 + N):
    pool.push({s: pcm[s][k * PUSH:(k + 1) * PUSH].tobytes() for s in range(S)})
torch.cuda.synchronize(); tot = time.perf_counter() - t0
# pure device time of the chunk-step graph
g = pool.pool._graph
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gms = None
if g is not None:
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); e1.synchronize()
    gms = e0.elapsed_time(e1) / 10
out = {k: round(v / N * 1e3, 3) for k, v in T.items()}
out["push total (instrumented, ms)"] = round(tot / N * 1e3, 3)
out["host rest (PCM decode, ring index ops, greedy fold, Python) ms"] = round((tot - sum(T.values())) / N * 1e3, 3)
out["chunk-step CUDA graph, device time (ms)"] = gms
out["launches in the chunk-step graph"] = pool.pool._graph_launches
print(json.dumps(out))
