"""CPU: the host logic of ``StreamPool.push`` (per-stream sample carry-over with in-place renormalisation, device feature
ring, 67/64/3 windowing, batched rounds, incremental greedy history) driven by a stand-in engine built from the ORACLE
(fbank + chunk forward on torch-CPU).  Three interleaved streams — one of them the utterance frozen from the reference's
``MASRPredictor.predict_stream`` (tests/golden/predictor_golden.json) — must return, push by push, what one
``predict_stream`` per stream returns.  (The oracle is the stand-in here, i.e. test infrastructure; the product classes
under test are StreamPool and its ring/windowing logic.)"""
import numpy as np
import pytest
import torch

from conftest import make_audio, synth_weights
from masr_b200 import stream_pool as sp, synth
from masr_b200.engine import subsampled_len
from masr_b200.predict import CACHED_FEATURE_NUM, DECODING_WINDOW, chunk_starts
from oracle import conformer as oc, ctc as octc, fbank as ob


class OracleEngine:
    """The two engine entry points StreamPool uses, on the CPU."""
    device = torch.device("cpu")

    def __init__(self):
        self.last_gain = None

    def fbank(self, waves, use_db=True, target_db=-20.0):
        feats, frames, gains, status = [], [], [], []
        for w in waves:
            try:
                x, g = ob.normalize_gain(np.asarray(w, np.float32).copy(), target_db) if use_db else (w, np.float32(1))
                status.append(0)
            except ValueError:                     # the CUDA front-end reports > 300 dB as a per-utterance status flag
                x, g = np.zeros(0, np.float32), np.float32(1)
                status.append(1)
            f = ob.kaldi_fbank(ob.to_int16(x)) if x.shape[0] >= 400 else np.zeros((0, 80), np.float32)
            feats.append(f); frames.append(f.shape[0]); gains.append(g)
        Fmax = max(1, max(frames))
        out = torch.zeros(len(waves), Fmax, 80)
        for i, f in enumerate(feats):
            out[i, :f.shape[0]] = torch.from_numpy(f)
        self.last_gain = torch.tensor(gains, dtype=torch.float32)
        return out, frames, torch.tensor(status, dtype=torch.int32)


class OraclePool:
    """``pool.step`` semantics of the CUDA pools (ids / max-prob per slot, valid counts) from the oracle chunk forward."""

    def __init__(self, sd, n_slots):
        self.sd, self.cfg, self.S = sd, oc.ConformerConfig(), n_slots
        self.st = [oc.ChunkState() for _ in range(n_slots)]

    def reset(self, slot):
        self.st[slot] = oc.ChunkState()

    def step(self, feats, nframes):
        ids = torch.zeros(self.S, 16, dtype=torch.int32)
        maxp = torch.zeros(self.S, 16)
        tout = [subsampled_len(int(n)) for n in nframes]
        for s, n in enumerate(nframes):
            if tout[s]:
                with torch.no_grad():
                    probs = oc.get_encoder_out_chunk(self.sd, self.cfg, feats[s:s + 1, :n], self.st[s], -16)[0]
                ids[s, :tout[s]] = probs.argmax(1).to(torch.int32)
                maxp[s, :tout[s]] = probs.max(1).values
        return ids, maxp, tout


def oracle_predict_stream(sd, pcm, push, vocab):
    st, gs, cfg = oc.ChunkState(), octc.GreedyStream(), oc.ConformerConfig()
    remained, cached, out = None, None, []
    for s in range(0, len(pcm), push):
        is_end = s + push >= len(pcm)
        new = ob.pcm_bytes_to_float32(pcm[s:s + push].tobytes())
        remained = new if remained is None else np.concatenate([remained, new])
        x, _ = ob.normalize_gain(remained.copy())
        feat = ob.kaldi_fbank(ob.to_int16(x))
        cached = feat if cached is None else np.concatenate([cached, feat], axis=0)
        remained = x[160 * feat.shape[0]:]
        starts = chunk_starts(cached.shape[0], is_end)
        if not starts:
            out.append(None)
            continue
        res, end = None, None
        for cur in starts:
            end = min(cur + DECODING_WINDOW, cached.shape[0])
            with torch.no_grad():
                probs = oc.get_encoder_out_chunk(sd, cfg, torch.from_numpy(cached[cur:end])[None], st, -16)[0].numpy()
            res = gs.push(probs, vocab)
        cached = cached[end - CACHED_FEATURE_NUM:]
        out.append({"text": res[1], "score": res[0]})
    return out


def test_stream_pool_host_logic_matches_predict_stream(monkeypatch, predictor_golden):
    g = predictor_golden
    sd = synth.to_torch(synth_weights(g["wseed"]))
    vocab = synth.vocabulary()
    S = 4
    monkeypatch.setattr(sp, "make_pool", lambda eng, n, max_frames=3000: OraclePool(sd, n))
    monkeypatch.setattr(sp.StreamPool, "RING", 256)            # small ring: the wrap-around path is exercised
    pool = sp.StreamPool(OracleEngine(), vocab, n_slots=S)
    audios = [make_audio(g["kind"], g["aseed"], g["samples"]), make_audio("noise", 95, 16000 * 2 + 3000), make_audio("speech", 96, 16000 * 4)]
    pcms = [(np.clip(a, -1, 1) * 32767).astype("<i2") for a in audios]
    push = g["push"]
    want = [oracle_predict_stream(sd, p, push, vocab) for p in pcms]
    for r, w in zip(want[0], g["pushes_pcm"]):                # the stand-in itself reproduces the reference's frozen pushes
        assert (r is None) == (w is None) and (r is None or (r["text"] == w["text"] and abs(r["score"] - w["score"]) < 1e-3))
    got = [[] for _ in pcms]
    npush = [len(range(0, len(p), push)) for p in pcms]
    for k in range(max(npush)):
        mid = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(len(pcms)) if k < npush[i] - 1}
        last = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(len(pcms)) if k == npush[i] - 1}
        for grp, is_end in ((mid, False), (last, True)):
            if grp:
                out = pool.push(grp, is_end=is_end)
                for i in grp:
                    got[i].append(out[i])
    for i in range(len(pcms)):
        assert len(got[i]) == len(want[i])
        for r, w in zip(got[i], want[i]):
            assert (r is None) == (w is None), (i, r, w)
            if r is not None:
                assert r["text"] == w["text"], (i, r, w)
                assert abs(r["score"] - w["score"]) < 1e-4
    # a slot can be reset and reused
    pool.reset_stream(1)
    out = pool.push({1: pcms[1][:push * 3].tobytes()}, is_end=True)
    ref = oracle_predict_stream(sd, pcms[1][:push * 3], push * 3, vocab)
    assert out[1]["text"] == ref[-1]["text"]


def test_stream_sessions_protocol(monkeypatch, predictor_golden):
    """infer_server.py's websocket contract over the pool: replies after every non-empty message, `end` suffix closes the
    utterance and frees the slot, no-resource answer when every slot is busy."""
    from masr_b200 import serve
    g = predictor_golden
    sd = synth.to_torch(synth_weights(g["wseed"]))
    vocab = synth.vocabulary()
    monkeypatch.setattr(sp, "make_pool", lambda eng, n, max_frames=3000: OraclePool(sd, n))
    sess = serve.StreamSessions(sp.StreamPool(OracleEngine(), vocab, n_slots=2))
    a, b = sess.open(), sess.open()
    assert a is not None and b is not None and a != b and sess.open() is None       # third connection: no resource
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = g["push"]
    msgs = [pcm[s:s + push].tobytes() for s in range(0, len(pcm), push)]
    msgs[-1] += b"end"
    want_text, texts = "", []
    for w in g["pushes_pcm"]:
        if w is not None:
            want_text = w["text"]
        texts.append(want_text)
    for k, m in enumerate(msgs):
        batch = {a: m}
        if k == 1:
            batch[b] = b""                                    # empty message: ignored, no reply
        rep = sess.feed(batch)
        assert set(rep) == {a}
        assert rep[a] == {"code": 0, "result": texts[k]}
    assert a not in sess.text and len(sess.free) == 1          # closed by `end`
    c = sess.open()
    assert c == a                                              # the freed slot is reused with a clean state
    rep = sess.feed({c: msgs[0], b: msgs[0]})
    assert rep[c] == rep[b] == {"code": 0, "result": texts[0]}
    with pytest.raises(KeyError):
        sess.feed({99: b"xx"})


def test_push_isolates_a_failing_slot_and_grows_the_ring(monkeypatch, predictor_golden):
    """a slot that cannot be decoded (gain above 300 dB, capacity, garbage input) must fail ALONE — its state
    untouched, the other slots of the same push decoded normally and identically to a push without the bad slot; a
    single message longer than the feature ring (about 10 s) is consumed like the reference consumes it."""
    from masr_b200 import serve
    g = predictor_golden
    sd = synth.to_torch(synth_weights(g["wseed"]))
    vocab = synth.vocabulary()
    monkeypatch.setattr(sp, "make_pool", lambda eng, n, max_frames=3000: OraclePool(sd, n))
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = g["push"]
    silent = np.full(8000, 1e-20, np.float32)                  # needs > 300 dB of gain -> ValueError in the reference

    clean = sp.StreamPool(OracleEngine(), vocab, n_slots=2)
    mixed = sp.StreamPool(OracleEngine(), vocab, n_slots=2)
    nmsg = len(range(0, len(pcm), push))
    for k in range(nmsg):
        m = pcm[k * push:(k + 1) * push].tobytes()
        want = clean.push({0: m}, is_end=k == nmsg - 1)
        if k == 1:
            with pytest.raises(sp.StreamSlotError) as ei:
                mixed.push({0: m, 1: silent}, is_end=False)
            assert set(ei.value.errors) == {1} and isinstance(ei.value.errors[1], ValueError)
            got = ei.value.results
        else:
            got = mixed.push({0: m, 1: silent}, is_end=k == nmsg - 1, on_error="return")
            assert set(mixed.last_errors) == {1}
        assert got == want                                     # the healthy slot is unaffected, push by push
        assert mixed.remained[1] is None and mixed.count[1] == 0 and mixed.toks[1] == []     # failing slot: state untouched
    # the serving layer fails only the offending session
    sess = serve.StreamSessions(sp.StreamPool(OracleEngine(), vocab, n_slots=2))
    a, b = sess.open(), sess.open()
    rep = sess.feed({a: pcm[:push].tobytes(), b: b"\x01"})     # odd byte count: not int16 PCM
    assert rep[b] == serve.FAILED and rep[a]["code"] == 0
    # one 12 s message (1198 frames > RING = 1024): same result as the single-stream oracle fed the same message
    long_pcm = (np.clip(make_audio("speech", 97, 16000 * 12), -1, 1) * 32767).astype("<i2")
    pool = sp.StreamPool(OracleEngine(), vocab, n_slots=2)
    first = pool.push({0: pcm[:push].tobytes()}, is_end=False)                     # slot 0 mid-utterance while the ring grows
    out = pool.push({1: long_pcm.tobytes()}, is_end=True)
    ref = oracle_predict_stream(sd, long_pcm, len(long_pcm), vocab)
    assert pool.RING >= 2048 and out[1]["text"] == ref[-1]["text"] and abs(out[1]["score"] - ref[-1]["score"]) < 1e-4
    rest = pool.push({0: pcm[push:2 * push].tobytes()}, is_end=False)
    c2 = sp.StreamPool(OracleEngine(), vocab, n_slots=1)
    assert c2.push({0: pcm[:push].tobytes()}) == first and c2.push({0: pcm[push:2 * push].tobytes()}) == rest
