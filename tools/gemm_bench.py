"""Dev tool: time masr_gemm_tc_f16x2 on the headline step's GEMM shapes (M = 32 x 248 frames) with CUDA events.
20 back-to-back launches captured in a CUDA graph (no host gaps), best of 5 replays, per MASR_TC_FLAGS value; prints us/launch and algorithmic TFLOP/s.  Not a bench value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from masr_b200 import _lib

_lib.load()
_lib.call("masr_check_device")
dev = torch.device("cuda", torch.cuda.current_device())
M = int(os.environ.get("GB_M", "7936"))
st = torch.cuda.current_stream().cuda_stream
REP = 20


def P(t):
    return None if t is None else t.data_ptr()


def pair(r, c):
    x = torch.randn(r, c, device=dev)
    h = torch.empty(r, c, dtype=torch.float16, device=dev)
    l = torch.empty_like(h)
    _lib.call("masr_split_f16", P(x), P(h), P(l), x.numel(), st)
    return h, l


# name, N, K, epilogue, fp32 out, pair out, residual
SHAPES = [("ffn_w1", 2048, 256, 1, False, True, False), ("ffn_w2", 256, 2048, 5, True, False, True),
          ("qkv", 768, 256, 0, True, True, False), ("out/pw2", 256, 256, 5, True, False, True),
          ("pw1_glu", 512, 256, 3, True, False, False), ("embed", 256, 4864, 4, True, False, False),
          ("ctc", 4233, 256, 0, True, False, False)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, N, K, epi, want_c, want_p, want_r in SHAPES:
    Ah, Al = pair(M, K)
    Wh, Wl = pair(N, K)
    b = torch.randn(N, device=dev)
    No = N // 2 if epi == 3 else N
    ldc = (No + 7) // 8 * 8
    C = torch.empty(M, ldc, device=dev) if want_c else None
    Ch = torch.empty(M, ldc, dtype=torch.float16, device=dev) if want_p else None
    Cl = torch.empty_like(Ch) if want_p else None
    R = torch.randn(M, ldc, device=dev) if want_r else None

    def run():
        _lib.call("masr_gemm_tc_f16x2", P(Ah), P(Al), K, P(Wh), P(Wl), P(b), P(R), ldc, P(C), P(Ch), P(Cl), ldc, M, N, K, epi,
                  0.5, st)

    for flags in os.environ.get("GB_FLAGS", "0,1,3").split(","):
        os.environ["MASR_TC_FLAGS"] = flags
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st = side.cuda_stream
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                for _ in range(REP):
                    run()
        st = torch.cuda.current_stream().cuda_stream
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        us = best / REP * 1e3
        print(f"{name:8s} N={N:5d} K={K:5d} flags={flags} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TF/s algorithmic "
              f"({6.0 * M * N * K / us / 1e6:7.1f} executed)", flush=True)
