"""Squeezeformer chunk (streaming) path: the oracle's ``get_encoder_out_chunk`` restatement reproduces what the reference's
``MASRPredictor.predict_stream`` returned (tests/golden/predictor_golden_squeezeformer.json, made by make_golden.py);
the CUDA stream pool reproduces the oracle chunk by chunk and the golden push by push (-m gpu)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, make_audio
from masr_b200 import synth
from masr_b200.predict import CACHED_FEATURE_NUM, DECODING_WINDOW, chunk_starts
from oracle import ctc as octc, fbank as ob, squeezeformer as osq


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "predictor_golden_squeezeformer.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def weights(golden):
    return synth.squeezeformer_state_dict(golden["wseed"], streaming=True)


def oracle_predict_stream(sd, cfg, pcm, push, vocab):
    """``MASRPredictor.predict_stream`` (predict.py:237-343) on top of the oracle: per push featurise the un-consumed
    samples (dB-normalised in place, so the tail keeps the gain), decode every complete 67-frame window."""
    st, gs = osq.ChunkState(), octc.GreedyStream()
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
                probs = osq.get_encoder_out_chunk(sd, cfg, torch.from_numpy(cached[cur:end])[None], st, -16)[0].numpy()
            res = gs.push(probs, vocab)
        cached = cached[end - CACHED_FEATURE_NUM:]
        out.append({"text": res[1], "score": res[0]})
    return out


def test_oracle_chunk_path_reproduces_reference_predict_stream(golden, weights):
    sd = synth.to_torch(weights)
    cfg = osq.SqueezeformerConfig(causal=True)
    x = make_audio(golden["kind"], golden["aseed"], golden["samples"])
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    got = oracle_predict_stream(sd, cfg, pcm, golden["push"], synth.vocabulary())
    assert len(got) == len(golden["pushes_pcm"])
    for r, w in zip(got, golden["pushes_pcm"]):
        assert (r is None) == (w is None)
        if r is not None:
            assert r["text"] == w["text"]
            assert abs(r["score"] - w["score"]) < 1e-3


def build_predictor(tmp, sd):
    import yaml
    from masr_b200.predict import MASRPredictor
    mp, vp = str(tmp / "sqz.pt"), str(tmp / "vocabulary.txt")
    torch.save(synth.to_torch(sd), mp)
    synth.write_vocabulary(vp)
    cfg = {"use_model": "squeezeformer", "streaming": True, "decoder": "ctc_greedy",
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp}}
    return MASRPredictor(configs=cfg, model_path=mp, use_gpu=True)


@pytest.mark.gpu
def test_gpu_predict_stream_matches_reference_golden(tmp_path, golden, weights):
    pred = build_predictor(tmp_path, weights)
    x = make_audio(golden["kind"], golden["aseed"], golden["samples"])
    whole = pred.predict(audio_data=x.copy())
    assert whole["text"] == golden["whole"]["text"] and abs(whole["score"] - golden["whole"]["score"]) < 1e-3
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = golden["push"]
    for rep in range(2):                                   # twice: reset_stream must restore a clean state
        pred.reset_stream()
        got = [pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)]
        assert len(got) == len(golden["pushes_pcm"])
        for r, w in zip(got, golden["pushes_pcm"]):
            assert (r is None) == (w is None), (rep, r, w)
            if r is not None:
                assert r["text"] == w["text"], (rep, r, w)
                assert abs(r["score"] - w["score"]) < 1e-3


@pytest.mark.gpu
def test_gpu_pool_matches_oracle_chunk_by_chunk(weights):
    """Three slots with different utterances (one joins late, one ends with a short chunk, one idles): per-frame ids are
    bit-exact and max-probabilities within 5e-5 of the oracle's chunk forward, slot by slot."""
    from masr_b200.squeezeformer import SqueezeformerEngine
    from masr_b200.stream_pool import SqueezeformerStreamPool
    eng = SqueezeformerEngine(weights, streaming=True)
    sd = synth.to_torch(weights)
    cfg = osq.SqueezeformerConfig(causal=True)
    feats = [ob.featurize(make_audio(k, seed, n)) for k, seed, n in
             [("speech", 40, 16000 * 3 + 2000), ("noise", 41, 16000 * 2 + 9000), ("speech", 42, 16000 * 2)]]
    S = 4
    pool = SqueezeformerStreamPool(eng, S, max_frames=400, keep_probs=True)     # + the chunk posteriors (InferencePredictor seam)
    states = [osq.ChunkState() for _ in feats]
    # chunk schedule per slot: list of (cur, end) windows like predict_stream with is_end at the end
    sched = []
    for f in feats:
        nf = f.shape[0]
        sched.append([(c, min(c + DECODING_WINDOW, nf)) for c in chunk_starts(nf, True)])
    delay = [0, 1, 0]                                       # slot 1 joins one round late
    rounds = max(len(s) + d for s, d in zip(sched, delay))
    batch = torch.zeros(S, DECODING_WINDOW, 80, device=eng.device)
    saw_short = False
    for r in range(rounds):
        nfr = [0] * S
        for i, f in enumerate(feats):
            k = r - delay[i]
            if 0 <= k < len(sched[i]):
                cur, end = sched[i][k]
                batch[i, :end - cur].copy_(torch.from_numpy(f[cur:end]))
                nfr[i] = end - cur
                saw_short |= (end - cur) < DECODING_WINDOW
        ids, maxp, tout = pool.step(batch, nfr)
        ids_h, mp_h = ids.cpu().numpy(), maxp.cpu().numpy()
        for i, f in enumerate(feats):
            if nfr[i] == 0:
                assert tout[i] == 0
                continue
            cur, end = sched[i][r - delay[i]]
            with torch.no_grad():
                probs = osq.get_encoder_out_chunk(sd, cfg, torch.from_numpy(f[cur:end])[None], states[i], -16)[0].numpy()
            assert tout[i] == probs.shape[0]
            assert np.array_equal(ids_h[i, :tout[i]], probs.argmax(1)), (r, i)
            assert np.abs(mp_h[i, :tout[i]] - probs.max(1)).max() < 5e-5, (r, i)
            got = pool.probs.view(S, -1, probs.shape[1])[i, :tout[i]].cpu().numpy()
            assert np.abs(got - probs).max() < 5e-5, (r, i)
    assert saw_short
    # a slot that decoded a short chunk must be reset before it is used again
    with pytest.raises(AssertionError):
        pool.step(batch, [67, 0, 0, 0])
    pool.reset(0)
    ids, maxp, tout = pool.step(batch, [67, 0, 0, 0])
    st = osq.ChunkState()
    with torch.no_grad():
        probs = osq.get_encoder_out_chunk(sd, cfg, batch[0:1].cpu(), st, -16)[0].numpy()
    assert np.array_equal(ids.cpu().numpy()[0, :16], probs.argmax(1))


@pytest.mark.gpu
def test_gpu_stream_pool_concurrent_streams_equal_single_stream(tmp_path, golden, weights):
    from masr_b200.stream_pool import StreamPool
    pred = build_predictor(tmp_path, weights)
    audios = [make_audio(golden["kind"], golden["aseed"], golden["samples"]), make_audio("noise", 93, 16000 * 2 + 4000),
              make_audio("speech", 94, 16000 * 3 + 1234)]
    pcms = [(np.clip(a, -1, 1) * 32767).astype("<i2") for a in audios]
    push = 8000
    want = []
    for pcm in pcms:
        pred.reset_stream()
        want.append([pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)])
    pred.reset_stream()
    pool = StreamPool(pred.predictor, synth.vocabulary(), n_slots=4, max_frames=600)
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
                assert abs(r["score"] - w["score"]) < 1e-3
