"""Dev tool: the kernel forms of tc_gemm (single CTA / cta_group::2 pair, with / without alternating epilogue groups, the
library default) on the same operands — outputs must be bit-identical (same products, same accumulation order per output
row) — then each timed (20 launches in a CUDA graph, best
of 5 replays, CUDA events).  Shapes: the headline step's GEMMs (M = 32 x 248 frames), ragged M / N, the CTC head, the conv-2
implicit GEMM.  Not a bench value."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from masr_b200 import _lib

_lib.load()
_lib.call("masr_check_device")
dev = torch.device("cuda", torch.cuda.current_device())
REP = 20
results = []


def P(t):
    return None if t is None else t.data_ptr()


def cur():
    return torch.cuda.current_stream().cuda_stream


def split(x):
    h = torch.empty(x.shape, dtype=torch.float16, device=dev)
    l = torch.empty_like(h)
    _lib.call("masr_split_f16", P(x), P(h), P(l), x.numel(), cur())
    return h, l


def time_graph(run):
    for _ in range(2):
        run(cur())
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            for _ in range(REP):
                run(side.cuda_stream)
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
    return best / REP * 1e3


MODES = (("single", "0", "5"), ("pair", "1", "5"), ("single+alt", "0", "13"), ("pair+alt", "1", "13"), ("default", None, None))


def set_mode(pair, flags):
    for k, v in (("MASR_TC_PAIR", pair), ("MASR_TC_FLAGS", flags)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def bits(t):
    return t.view(torch.int32) if t.dtype == torch.float32 else t.view(torch.int16) if t.dtype == torch.float16 else t


def ab(name, make_outputs, run, flops):
    """run(stream, outs) launches the op into `outs`.  Every kernel form (single CTA / cta_group::2 pair, with and without the
    alternating epilogue groups, and the library's default choice) must give bit-identical outputs; then each is timed."""
    outs, t = {}, {}
    for mode, pair, flags in MODES:
        set_mode(pair, flags)
        o = make_outputs()
        for x in o:
            if x is not None:
                x.fill_(-77 if x.dtype == torch.int32 else float("nan"))
        run(cur(), o)
        torch.cuda.synchronize()
        outs[mode] = o
    same = all(a is None or bool(torch.equal(bits(a), bits(b))) for m in outs for a, b in zip(outs["single"], outs[m]))
    for mode, pair, flags in MODES:
        set_mode(pair, flags)
        o = outs[mode]
        t[mode] = round(time_graph(lambda s: run(s, o)), 2)
    set_mode(None, None)
    r = {"op": name, "bit_identical": same, "us": t, "default_vs_single": round(t["single"] / t["default"], 3),
         "algorithmic_tflops_default": round(flops / t["default"] / 1e6, 1)}
    results.append(r)
    print(json.dumps(r), flush=True)


def gemm_case(name, M, N, K, epi, want_c, want_p, want_r, alpha=0.5):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.05
    Ah, Al = split(A)
    Wh, Wl = split(W)
    b = torch.randn(N, device=dev)
    No = N // 2 if epi == 3 else N
    ldc = (No + 7) // 8 * 8
    R = torch.randn(M, ldc, device=dev) if want_r else None

    def mk():
        return [torch.empty(M, ldc, device=dev) if want_c else None,
                torch.empty(M, ldc, dtype=torch.float16, device=dev) if want_p else None,
                torch.empty(M, ldc, dtype=torch.float16, device=dev) if want_p else None]

    def run(s, o):
        _lib.call("masr_gemm_tc_f16x2", P(Ah), P(Al), K, P(Wh), P(Wl), P(b), P(R), ldc, P(o[0]), P(o[1]), P(o[2]), ldc, M, N, K, epi,
                  alpha, s)

    ab(name, mk, run, 2.0 * M * N * K)      # (padding columns keep their fill pattern in both runs)


only = set(sys.argv[1:])


def want(n):
    return not only or n in only


M = 7936
if want("ffn_w1"):
    gemm_case("ffn_w1 M7936 N2048 K256 bias+SiLU pair-out", M, 2048, 256, 1, False, True, False)
if want("ffn_w2"):
    gemm_case("ffn_w2 M7936 N256 K2048 residual fp32", M, 256, 2048, 5, True, False, True)
if want("qkv"):
    gemm_case("qkv M7936 N768 K256 bias fp32+pair", M, 768, 256, 0, True, True, False)
if want("out"):
    gemm_case("out/pw2 M7936 N256 K256 residual", M, 256, 256, 5, True, False, True, alpha=1.0)
if want("glu"):
    gemm_case("pw1 M7936 N512 K256 GLU", M, 512, 256, 3, True, True, False)
if want("embed"):
    gemm_case("embed M7936 N256 K4864 scale", M, 256, 4864, 4, True, False, False)
if want("ragged"):
    gemm_case("ragged M8013 N264 K320 bias (odd row blocks, partial column tile)", 8013, 264, 320, 0, True, True, False)
    gemm_case("ragged M300 N4233 K256 bias (3 row blocks, vocabulary width)", 300, 4233, 256, 0, True, False, False)
    gemm_case("ragged M129 N256 K256 relu", 129, 256, 256, 2, True, True, False)

if want("ctc"):
    V, K = 4233, 256
    A = torch.randn(M, K, device=dev)
    W = torch.randn(V, K, device=dev) * 0.05
    Ah, Al = split(A)
    Wh, Wl = split(W)
    b = torch.randn(V, device=dev)
    groups = (V + 31) // 32
    ws = torch.empty(3 * groups * M, dtype=torch.float32, device=dev)

    def mk():
        return [torch.empty(M, dtype=torch.int32, device=dev), torch.empty(M, device=dev)]

    def run(s, o):
        _lib.call("masr_ctc_head_argmax_tc_f16x2", P(Ah), P(Al), K, P(Wh), P(Wl), P(b), M, V, K, P(ws), ws.numel() * 4, P(o[0]), P(o[1]), s)
    ab("ctc head M7936 V4233 K256 (argmax + max-prob)", mk, run, 2.0 * M * V * K)

if want("conv"):
    B, F1, C = 32, 995, 256
    TH = (F1 + 1) // 2
    T2 = (F1 - 3) // 2 + 1
    planes = torch.randn(4, B, TH, 20, C, device=dev).clamp_(min=0)
    ch, cl = split(planes)
    W = torch.randn(C, 9 * C, device=dev) * 0.02
    Wh, Wl = split(W)
    b = torch.randn(C, device=dev)
    rows = B * T2 * 19

    def mk():
        return [torch.empty(rows, C, device=dev), torch.empty(rows, C, dtype=torch.float16, device=dev),
                torch.empty(rows, C, dtype=torch.float16, device=dev)]

    def run(s, o):
        _lib.call("masr_conv2_tc_f16x2", P(ch), P(cl), P(Wh), P(Wl), P(b), P(o[0]), P(o[1]), P(o[2]), B, F1, T2, C, s)
    ab(f"conv2 implicit GEMM B{B} T2={T2} (rows {rows}) N256 K2304", mk, run, 2.0 * rows * C * 9 * C)
    # odd number of 6-row time tiles and a short utterance
    for Bs, F1s in ((3, 61), (2, 15)):
        THs, T2s = (F1s + 1) // 2, (F1s - 3) // 2 + 1
        pl = torch.randn(4, Bs, THs, 20, C, device=dev)
        h2, l2 = split(pl)
        rws = Bs * T2s * 19

        def mk2(rws=rws):
            return [torch.empty(rws, C, device=dev), torch.empty(rws, C, dtype=torch.float16, device=dev),
                    torch.empty(rws, C, dtype=torch.float16, device=dev)]

        def run2(s, o, h2=h2, l2=l2, Bs=Bs, F1s=F1s, T2s=T2s):
            _lib.call("masr_conv2_tc_f16x2", P(h2), P(l2), P(Wh), P(Wl), P(b), P(o[0]), P(o[1]), P(o[2]), Bs, F1s, T2s, C, s)
        ab(f"conv2 B{Bs} T2={T2s}", mk2, run2, 2.0 * rws * C * 9 * C)

ok = all(r["bit_identical"] for r in results)
print(json.dumps({"all_bit_identical": ok, "cases": len(results)}))
sys.exit(0 if ok else 1)
