"""Dev tool: where the end-to-end time of ConformerEngine.transcribe goes (host staging, H2D, graph replay, D2H, text)."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from masr_b200 import synth, _lib
from masr_b200.engine import ConformerEngine
print("host cores", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
eng = ConformerEngine(synth.conformer_state_dict(0), streaming=True)
waves = [synth.noise_audio(1000 + i, 160000) for i in range(32)]
for _ in range(3): eng.transcribe(waves)
torch.cuda.synchronize()
def T(): torch.cuda.synchronize(); return time.perf_counter()
N = 20
t0 = T()
for _ in range(N): eng.transcribe(waves)
t1 = T(); print("transcribe total ms", (t1 - t0) * 1e3 / N)
step = eng.prepare_resident(waves)
t0 = T()
for _ in range(N): step()
t1 = T(); print("graph replay only ms (L2 warm)", (t1 - t0) * 1e3 / N)
pin = torch.empty(32 * 160000, dtype=torch.float32, pin_memory=True); hv = pin.numpy()
d = torch.empty(32 * 160000, device="cuda")
t0 = time.perf_counter()
for _ in range(N):
    for i, w in enumerate(waves): hv[i * 160000:(i + 1) * 160000] = w
t1 = time.perf_counter(); print("numpy staging memcpy ms", (t1 - t0) * 1e3 / N)
t0 = T()
for _ in range(N): d.copy_(pin, non_blocking=True)
t1 = T(); print("H2D 20MB single copy ms", (t1 - t0) * 1e3 / N)
ptrs = (C.c_void_p * 32)(*[w.ctypes.data for w in waves]); lens = (C.c_int64 * 32)(*[160000] * 32)
st = torch.cuda.current_stream().cuda_stream
for th in (1, 2, 4, 8, 16):
    t0 = T()
    for _ in range(N): _lib.call("masr_stage_waves_f32", ptrs, lens, 32, pin.data_ptr(), d.data_ptr(), th, st)
    t1 = T(); print(f"native stage+H2D threads={th} ms", (t1 - t0) * 1e3 / N)
ws = step()
t0 = T()
for _ in range(N):
    a = ws["out_pack"].cpu()
t1 = T(); print("packed D2H (.cpu) ms", (t1 - t0) * 1e3 / N)
from masr_b200.text import ids_to_text
vocab = synth.vocabulary()
res = eng.transcribe(waves)
t0 = time.perf_counter()
for _ in range(N): [ids_to_text(t, vocab) for t in res.tokens]
t1 = time.perf_counter(); print("ids_to_text x32 ms", (t1 - t0) * 1e3 / N)
