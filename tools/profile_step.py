"""Profiling target (dev tool): one steady-state pass of the hot path at the headline size between
cudaProfilerStart/Stop, for `ncu --profile-from-start off`.  Numbers printed under a profiler are
never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from masr_b200 import synth
from masr_b200.engine import ConformerEngine

B, N = int(os.environ.get("PROF_B", "32")), 160000
eng = ConformerEngine(synth.conformer_state_dict(0), streaming=True)
waves = [synth.noise_audio(1000 + i, N) for i in range(B)]
dev = eng.device
offs = torch.tensor(np.arange(B + 1, dtype=np.int64) * N, device=dev)
wave_dev = torch.from_numpy(np.concatenate(waves)).to(dev)


def step():
    feats, frames, status = eng.fbank(None, True, -20.0, wave_dev=wave_dev, offsets_dev=offs, lengths=[N] * B)
    enc, tl, T, ws = eng.encode(feats, frames)
    eng.ctc_greedy(enc, tl, T, ws)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step,", eng.launches // 3, "launches per step")
