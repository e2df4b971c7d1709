"""DeepSpeech2: oracle pinned to the reference's frozen outputs (CPU) and the CUDA engine against oracle + golden (GPU)."""
import numpy as np
import pytest
import torch

from conftest import load_npz, make_audio
from masr_b200 import synth
from oracle import ctc as octc, deepspeech2 as od, fbank as ob

_W = {}


def weights(seed, streaming):
    if (seed, streaming) not in _W:
        _W[seed, streaming] = synth.deepspeech2_state_dict(seed, streaming=streaming)
    return _W[seed, streaming]


def test_oracle_matches_reference_golden():
    z, meta = load_npz("deepspeech2_golden.npz")
    vocab = synth.vocabulary()
    for m in meta:
        sd = synth.to_torch(weights(m["wseed"], m["streaming"]))
        cfg = od.DS2Config(bidirectional=not m["streaming"])
        feat = torch.from_numpy(z[m["name"] + "/feat"])[None]
        with torch.no_grad():
            probs, _ = od.get_encoder_out(sd, cfg, feat)
        probs = probs.numpy()
        assert np.array_equal(probs.argmax(1), z[m["name"] + "/ids"])
        got = np.take_along_axis(probs, z[m["name"] + "/top_i"].astype(np.int64), axis=1)
        assert np.abs(got - z[m["name"] + "/top_p"]).max() < 1e-5
        score, text, _ = octc.greedy_decode(probs, vocab)
        assert text == m["text"] and abs(score - m["score"]) < 1e-3


def test_oracle_chunked_equals_whole_for_forward_lstm():
    """Carrying (h, c) across chunks reproduces the whole-utterance LSTM stack on the same subsampled frames."""
    sd = synth.to_torch(weights(0, True))
    cfg = od.DS2Config()
    feat = torch.from_numpy(ob.featurize(make_audio("speech", 5, 16000 * 2)))[None]
    st = None
    outs = []
    with torch.no_grad():
        for cur in range(0, feat.shape[1] - 67 + 1, 64):
            p, st = od.get_encoder_out(sd, cfg, feat[:, cur:cur + 67], st)
            outs.append(p)
        whole, _ = od.get_encoder_out(sd, cfg, feat[:, :64 * len(outs) + 3])
    assert torch.cat(outs).shape == whole.shape
    assert (torch.cat(outs) - whole).abs().max().item() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("streaming,wseed", [(True, 0), (False, 1)])
def test_gpu_engine_matches_oracle_and_golden(streaming, wseed):
    from masr_b200.deepspeech2 import DeepSpeech2Engine
    eng = DeepSpeech2Engine(weights(wseed, streaming), streaming=streaming)
    sd = synth.to_torch(weights(wseed, streaming))
    cfg = od.DS2Config(bidirectional=not streaming)
    vocab = synth.vocabulary()
    z, meta = load_npz("deepspeech2_golden.npz")
    for m in meta:
        if m["streaming"] != streaming:
            continue
        feat = z[m["name"] + "/feat"]
        res = eng.transcribe_features(torch.from_numpy(feat)[None].to(eng.device), [feat.shape[0]], None, return_frames=True)
        assert np.array_equal(res.frame_ids[0, :res.frame_lens[0]], z[m["name"] + "/ids"])
        assert "".join(vocab[i] for i in res.tokens[0]).replace("<space>", " ") == m["text"]
        assert abs(res.scores[0] - m["score"]) < 1e-3
        probs = eng.posteriors(feat[None], [feat.shape[0]])[0]
        got = np.take_along_axis(probs, z[m["name"] + "/top_i"].astype(np.int64), axis=1)
        assert np.abs(got - z[m["name"] + "/top_p"]).max() < 5e-5
    lens = [16000 * 2 + 17, 9000, 16000 + 320, 400 + 160 * 30]
    waves = [make_audio("speech" if i % 2 == 0 else "noise", 90 + i, n) for i, n in enumerate(lens)]
    res = eng.transcribe(waves, return_frames=True)
    for i, w in enumerate(waves):
        f = torch.from_numpy(ob.featurize(w.copy()))
        with torch.no_grad():
            probs, _ = od.get_encoder_out(sd, cfg, f[None])
        probs = probs.numpy()
        n = res.frame_lens[i]
        assert n == probs.shape[0]
        assert np.array_equal(probs.argmax(1), res.frame_ids[i, :n]), i
        score, text, toks = octc.greedy_decode(probs, vocab)
        assert toks == res.tokens[i] and abs(score - res.scores[i]) < 1e-3
    if streaming:
        feat = torch.from_numpy(ob.featurize(make_audio("speech", 8, 16000 * 2)))
        fd = feat.to(eng.device)
        st_o, st_g = None, eng.new_stream()
        for cur in range(0, feat.shape[0] - 67 + 1, 64):
            with torch.no_grad():
                pm, st_o = od.get_encoder_out(sd, cfg, feat[None, cur:cur + 67], st_o)
            ids, maxp, probs = eng.encode_chunk(fd[cur:cur + 67], st_g, want_probs=True)
            assert np.abs(probs.cpu().numpy() - pm.numpy()).max() < 5e-5
            assert np.array_equal(ids.cpu().numpy(), pm.numpy().argmax(1))


def test_oracle_chunk_path_reproduces_reference_predict_stream():
    """The oracle's chunked forward (LSTM state carried across 67-frame windows) + greedy history reproduces what the
    reference's real ``MASRPredictor.predict_stream`` returned push by push (tests/golden/predictor_golden_deepspeech2.json,
    made by make_golden.py)."""
    import json
    import os
    from conftest import GOLDEN
    from masr_b200.predict import CACHED_FEATURE_NUM, DECODING_WINDOW, chunk_starts
    with open(os.path.join(GOLDEN, "predictor_golden_deepspeech2.json"), encoding="utf-8") as f:
        g = json.load(f)
    sd = synth.to_torch(weights(g["wseed"], True))
    cfg = od.DS2Config(bidirectional=False)
    vocab = synth.vocabulary()
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    with torch.no_grad():
        probs, _ = od.get_encoder_out(sd, cfg, torch.from_numpy(ob.featurize(x.copy()))[None])
    score, text, _ = octc.greedy_decode(probs.numpy(), vocab)
    assert text == g["whole"]["text"] and abs(score - g["whole"]["score"]) < 1e-3
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = g["push"]
    state, gs = None, octc.GreedyStream()
    remained, cached, got = None, None, []
    for s in range(0, len(pcm), push):
        is_end = s + push >= len(pcm)
        new = ob.pcm_bytes_to_float32(pcm[s:s + push].tobytes())
        remained = new if remained is None else np.concatenate([remained, new])
        xn, _ = ob.normalize_gain(remained.copy())
        feat = ob.kaldi_fbank(ob.to_int16(xn))
        cached = feat if cached is None else np.concatenate([cached, feat], axis=0)
        remained = xn[160 * feat.shape[0]:]
        starts = chunk_starts(cached.shape[0], is_end)
        if not starts:
            got.append(None)
            continue
        res, end = None, None
        for cur in starts:
            end = min(cur + DECODING_WINDOW, cached.shape[0])
            with torch.no_grad():
                pr, state = od.get_encoder_out(sd, cfg, torch.from_numpy(cached[cur:end])[None], state)
            res = gs.push(pr.numpy(), vocab)
        cached = cached[end - CACHED_FEATURE_NUM:]
        got.append({"text": res[1], "score": res[0]})
    assert len(got) == len(g["pushes_pcm"])
    for r, w in zip(got, g["pushes_pcm"]):
        assert (r is None) == (w is None)
        if r is not None:
            assert r["text"] == w["text"]
            assert abs(r["score"] - w["score"]) < 1e-3
