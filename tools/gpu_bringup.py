"""Bring-up diagnostics (dev tool, GPU box): every kernel against the CPU oracle, printing max
abs/rel differences without stopping at the first mismatch.  Not part of the product."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from masr_b200 import synth, _lib
from masr_b200.engine import ConformerEngine, _p
from masr_b200._lib import call
from oracle import conformer as oc, fbank as ob, ctc as octc

torch.manual_seed(0)
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream


def report(name, a, b):
    a = torch.as_tensor(a).float().cpu(); b = torch.as_tensor(b).float().cpu()
    d = (a - b).abs()
    print(f"[{name}] shape {tuple(a.shape)} max_abs {d.max().item():.3e} mean_abs {d.mean().item():.3e} ref_absmax {b.abs().max().item():.3e} nan {int(torch.isnan(a).sum())}", flush=True)


def section(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


def t_gemm():
    for (M, N, K) in [(100, 256, 256), (333, 2048, 256), (77, 256, 2048), (500, 4233, 256), (7, 768, 256)]:
        A = torch.randn(M, K); W = torch.randn(N, K) / K ** 0.5; b = torch.randn(N); R = torch.randn(M, N)
        Ad, Wd, bd, Rd = A.to(dev), W.to(dev), b.to(dev), R.to(dev)
        ldc = (N + 15) // 16 * 16
        for epi, name in [(0, "bias"), (1, "silu"), (2, "relu"), (4, "scale"), (5, "resid")]:
            C = torch.zeros(M, ldc, device=dev)
            call("masr_gemm_f32", _p(Ad), K, _p(Wd), _p(bd), _p(Rd), N, _p(C), ldc, M, N, K, epi, 0.5, st())
            ref = F.linear(A, W, b)
            if epi == 1: ref = F.silu(ref)
            if epi == 2: ref = F.relu(ref)
            if epi == 4: ref = ref * 0.5
            if epi == 5: ref = R + 0.5 * ref
            report(f"gemm {M}x{N}x{K} {name}", C[:, :N], ref)
        if N % 4 == 0:
            Wi = torch.stack([W[:N // 2], W[N // 2:]], 1).reshape(N, K).to(dev); bi = torch.stack([b[:N // 2], b[N // 2:]], 1).reshape(N).to(dev)
            C = torch.zeros(M, N // 2, device=dev)
            call("masr_gemm_f32", _p(Ad), K, _p(Wi), _p(bi), None, 0, _p(C), N // 2, M, N, K, 3, 1.0, st())
            report(f"gemm {M}x{N}x{K} glu", C, F.glu(F.linear(A, W, b), dim=1))


def t_ln():
    x = torch.randn(1000, 256) * 3 + 1; g = torch.randn(256); b = torch.randn(256)
    y = torch.empty(1000, 256, device=dev)
    xd, gd, bd = x.to(dev), g.to(dev), b.to(dev)
    call("masr_layernorm_f32", _p(xd), 256, _p(gd), _p(bd), _p(y), 256, 1000, 256, 1e-5, st())
    report("layernorm", y, F.layer_norm(x, (256,), g, b, 1e-5))


def t_fbank(eng):
    waves = [synth.noise_audio(0, 16000 * 3), synth.speechlike_audio(1, 16000 * 2 + 123), synth.speechlike_audio(2, 5000), synth.noise_audio(3, 399), (synth.speechlike_audio(4, 30000) * 30)]
    feats, frames, status = eng.fbank(waves)
    print("frames", frames, "status", status.cpu().tolist())
    for i, w in enumerate(waves):
        ref = ob.featurize(w.copy())
        report(f"fbank utt{i} n={len(w)}", feats[i, :frames[i]], ref)
    feats2, _, _ = eng.fbank(waves, use_db_normalization=False)
    for i, w in enumerate(waves[:2]):
        report(f"fbank nonorm utt{i}", feats2[i, :frames[i]], ob.featurize(w.copy(), use_db_normalization=False))


def t_encoder(eng, sd, cfg):
    waves = [synth.speechlike_audio(10, 16000 * 4), synth.noise_audio(11, 16000 * 2 + 77), synth.speechlike_audio(12, 16000 * 3 + 5)]
    feats_h = [torch.from_numpy(ob.featurize(w.copy())) for w in waves]
    frames = [f.shape[0] for f in feats_h]
    Fmax = max(frames)
    fd = torch.zeros(len(waves), Fmax, 80)
    for i, f in enumerate(feats_h): fd[i, :f.shape[0]] = f
    enc, tl, T, ws = eng.encode(fd.to(dev), frames)
    torch.cuda.synchronize()
    enc = enc.view(len(waves), T, -1).cpu()
    with torch.no_grad():
        for i, f in enumerate(feats_h):
            taps = {}
            ref = oc.encode(sd, cfg, f[None], taps)[0]
            report(f"encoder utt{i} T={tl[i]}", enc[i, :tl[i]], ref)
    # stage taps for utt 0 alone
    with torch.no_grad():
        f = feats_h[0]
        taps = {}
        oc.encode(sd, cfg, f[None], taps)
        x = oc.subsample(sd, cfg, f[None])
    res = eng.transcribe_features(fd.to(dev), frames, None, return_frames=True)
    with torch.no_grad():
        for i, f in enumerate(feats_h):
            probs = oc.get_encoder_out(sd, cfg, f[None])[0].numpy()
            ids, mp = octc.best_path(probs)
            same = np.array_equal(ids, res.frame_ids[i, :tl[i]])
            s, _, toks = octc.greedy_decode(probs, synth.vocabulary(eng.V))
            print(f"[greedy utt{i}] frame ids equal {same} tokens equal {toks == res.tokens[i]} score ref {s:.6f} got {res.scores[i]:.6f} ntok {len(toks)}")
    pr = eng.posteriors(fd.numpy(), frames)
    with torch.no_grad():
        for i, f in enumerate(feats_h):
            probs = oc.get_encoder_out(sd, cfg, f[None])[0]
            report(f"posterior utt{i}", pr[i, :tl[i]], probs)


def t_e2e(eng, sd, cfg):
    waves = [synth.speechlike_audio(20 + i, 16000 * 2 + 1000 * i) for i in range(4)]
    res = eng.transcribe(waves, return_frames=True)
    vocab = synth.vocabulary(eng.V)
    with torch.no_grad():
        for i, w in enumerate(waves):
            f = torch.from_numpy(ob.featurize(w.copy()))
            probs = oc.get_encoder_out(sd, cfg, f[None])[0].numpy()
            ids, _ = octc.best_path(probs)
            s, txt, toks = octc.greedy_decode(probs, vocab)
            n = res.frame_lens[i]
            print(f"[e2e utt{i}] frames {n} ids equal {np.array_equal(ids, res.frame_ids[i, :n])} mism {int((ids != res.frame_ids[i, :n]).sum())} tokens equal {toks == res.tokens[i]} score {s:.5f} vs {res.scores[i]:.5f}")


def t_bench(eng):
    waves = [synth.noise_audio(100 + i, 160000) for i in range(32)]
    for _ in range(2): eng.transcribe(waves)
    torch.cuda.synchronize()
    t0 = time.time(); n = 3
    for _ in range(n): eng.transcribe(waves)
    torch.cuda.synchronize(); dt = (time.time() - t0) / n
    print(f"[bench] 32x10s e2e {dt*1e3:.2f} ms/step -> {320/dt:.0f} audio-s/s")
    feats, frames, status = eng.fbank(waves)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.transcribe_features(feats, frames)
    e1.record(); torch.cuda.synchronize()
    print(f"[bench] encoder+ctc from resident feats {e0.elapsed_time(e1)/n:.2f} ms")
    e0.record()
    for _ in range(10): eng.fbank(waves)
    e1.record(); torch.cuda.synchronize()
    print(f"[bench] fbank incl H2D {e0.elapsed_time(e1)/10:.3f} ms")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", _lib.load().masr_abi_version())
    section(t_gemm); section(t_ln)
    V = 4233
    sdn = synth.conformer_state_dict(0, V)
    sd = synth.to_torch(sdn)
    cfg = oc.ConformerConfig()
    eng = ConformerEngine(sdn, streaming=True)
    section(lambda: t_fbank(eng))
    section(lambda: t_encoder(eng, sd, cfg))
    section(lambda: t_e2e(eng, sd, cfg))
    section(lambda: t_bench(eng))
