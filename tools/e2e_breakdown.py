"""Dev tool: where the end-to-end time goes (host staging, H2D, graph replay, D2H)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from masr_b200 import synth
from masr_b200.engine import ConformerEngine
eng = ConformerEngine(synth.conformer_state_dict(0), streaming=True)
waves = [synth.noise_audio(1000 + i, 160000) for i in range(32)]
for _ in range(3): eng.transcribe(waves)
torch.cuda.synchronize()
def T(): torch.cuda.synchronize(); return time.perf_counter()
t0 = T()
for _ in range(10): eng.transcribe(waves)
t1 = T(); print("transcribe total ms", (t1 - t0) * 100)
step = eng.prepare_resident(waves)
t0 = T()
for _ in range(10): step()
t1 = T(); print("graph replay only ms", (t1 - t0) * 100)
pin = torch.empty(32 * 160000, dtype=torch.float32, pin_memory=True); hv = pin.numpy()
t0 = time.perf_counter()
for _ in range(10):
    for i, w in enumerate(waves): hv[i * 160000:(i + 1) * 160000] = w
t1 = time.perf_counter(); print("staging memcpy ms", (t1 - t0) * 100)
d = torch.empty(32 * 160000, device="cuda")
t0 = T()
for _ in range(10): d.copy_(pin, non_blocking=True)
t1 = T(); print("H2D 20MB ms", (t1 - t0) * 100)
t0 = time.perf_counter()
for _ in range(10): xs = [w.astype(np.float32) for w in waves]
t1 = time.perf_counter(); print("astype copies ms", (t1 - t0) * 100)
ws = step()
t0 = T()
for _ in range(10):
    a = ws["tokens"].cpu(); b = ws["ntok"].cpu(); c = ws["psum"].cpu(); e = ws["pcount"].cpu()
t1 = T(); print("4x D2H ms", (t1 - t0) * 100)
