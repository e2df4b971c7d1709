"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and against the
reference's frozen outputs (tests/golden).  Bars (BASELINE.json north_star):
  * integer work — int16 quantisation (via its effect), per-frame argmax ids, collapsed token ids: bit-exact;
  * floating point — fbank within FBANK_TOL (log-mel domain), encoder output within ENC_TOL, CTC
    posteriors within PROB_TOL, score within SCORE_TOL.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import make_audio, synth_weights
from masr_b200 import synth
from oracle import conformer as oc, ctc as octc, fbank as ob

pytestmark = pytest.mark.gpu

FBANK_TOL = 2e-3     # abs, log-mel (values ~5..26); independent float32 FFTs differ by ~1e-4..5e-4
ENC_TOL = 1e-4       # abs on LayerNorm-ed encoder output (|x| ~ 4)
PROB_TOL = 2e-5      # abs on softmax posteriors
SCORE_TOL = 1e-3     # reference score is 0..100


def test_fbank_vs_golden_and_oracle(gpu_engines, fbank_golden):
    eng = gpu_engines()
    z, meta = fbank_golden
    waves = [make_audio(m["kind"], m["seed"], m["samples"], m["scale"]) for m in meta]
    waves.append(synth.noise_audio(9, 399))            # shorter than one frame -> 0 frames
    feats, frames, status = eng.fbank(waves)
    assert frames[:-1] == [z[m["name"] + "/feat"].shape[0] for m in meta] and frames[-1] == 0
    assert status.cpu().tolist() == [0] * len(waves)
    f = feats.cpu().numpy()
    for i, m in enumerate(meta):
        ref = z[m["name"] + "/feat"]
        assert np.abs(f[i, :frames[i]] - ref).max() < FBANK_TOL, m["name"]
        assert np.all(f[i, frames[i]:] == 0)           # padded frames are deterministic zeros
    for i, w in enumerate(waves[:-1]):
        assert np.abs(f[i, :frames[i]] - ob.featurize(w.copy())).max() < FBANK_TOL


def test_fbank_gain_matches_numpy_chain(gpu_engines):
    eng = gpu_engines()
    waves = [make_audio("speech", s, 9000 + 777 * s, sc) for s, sc in [(1, 1.0), (2, 0.01), (3, 25.0)]]
    waves.append(np.zeros(1000, np.float32))           # all-zero audio: mean square 0 -> 1 (audio.py:526)
    eng.fbank(waves)
    gains = eng.last_gain.cpu().numpy()
    for g, w in zip(gains, waves):
        _, ref = ob.normalize_gain(w)
        assert abs(g / ref - 1) < 2e-6       # one float32 ulp of 10*log10(ms) moves the gain by ~1e-6
    # silence below 10^-32 mean square needs > 300 dB: the reference raises ValueError (audio.py:301)
    _, _, status = eng.fbank([np.full(2000, 1e-20, np.float32)])
    assert status.cpu().tolist() == [1]
    with pytest.raises(ValueError):
        ob.normalize_gain(np.full(2000, 1e-20, np.float32))


def test_encoder_golden(gpu_engines, conformer_golden):
    z, meta = conformer_golden
    vocab = synth.vocabulary()
    for m in meta:
        eng = gpu_engines(m["wseed"], m["streaming"])
        feat = z[m["name"] + "/feat"]
        fd = torch.from_numpy(feat)[None].to(eng.device)
        enc, tl, T, ws = eng.encode(fd, [feat.shape[0]])
        assert np.abs(enc.cpu().numpy() - z[m["name"] + "/enc"]).max() < ENC_TOL
        res = eng.transcribe_features(fd, [feat.shape[0]], None, return_frames=True)
        assert np.array_equal(res.frame_ids[0, :tl[0]], z[m["name"] + "/ids"])            # bit-exact ids
        assert "".join(vocab[i] for i in res.tokens[0]).replace("<space>", " ") == m["text"]
        assert abs(res.scores[0] - m["score"]) < SCORE_TOL
        probs = eng.posteriors(feat[None], [feat.shape[0]])[0]
        got = np.take_along_axis(probs, z[m["name"] + "/top_i"].astype(np.int64), axis=1)
        assert np.abs(got - z[m["name"] + "/top_p"]).max() < PROB_TOL
        assert np.abs(probs.sum(1) - 1).max() < 1e-5


@pytest.mark.parametrize("streaming,wseed", [(True, 0), (False, 1)])
def test_ragged_batch_equals_single_utterance_oracle(gpu_engines, streaming, wseed):
    """B=1 API semantics for every row of a padded batch, incl. the non-causal model where the
    reference's own padded batch path leaks padding into valid frames."""
    eng = gpu_engines(wseed, streaming)
    sd = synth.to_torch(synth_weights(wseed))
    cfg = oc.ConformerConfig(causal=streaming)
    lens = [16000 * 3 + 17, 9000, 16000 * 2, 400 + 160 * 6, 16000 * 4]
    waves = [make_audio("speech" if i % 2 == 0 else "noise", 40 + i, n) for i, n in enumerate(lens)]
    res = eng.transcribe(waves, return_frames=True)
    vocab = synth.vocabulary()
    mism = 0
    for i, w in enumerate(waves):
        f = torch.from_numpy(ob.featurize(w.copy()))
        with torch.no_grad():
            probs = oc.get_encoder_out(sd, cfg, f[None])[0].numpy()
        ids, _ = octc.best_path(probs)
        n = res.frame_lens[i]
        assert n == len(ids)
        mism += int((ids != res.frame_ids[i, :n]).sum())
        score, text, toks = octc.greedy_decode(probs, vocab)
        assert toks == res.tokens[i]
        assert abs(score - res.scores[i]) < SCORE_TOL
    assert mism == 0


def test_batch_is_permutation_and_batchsize_invariant(gpu_engines):
    """Size-independent properties at a larger size: results do not depend on batch composition."""
    eng = gpu_engines()
    waves = [synth.noise_audio(200 + i, 160000 if i % 3 else 120000 + 1000 * i) for i in range(12)]
    full = eng.transcribe(waves, return_frames=True)
    perm = [7, 2, 11, 0, 5, 9, 1, 3, 10, 4, 8, 6]
    shuf = eng.transcribe([waves[i] for i in perm], return_frames=True)
    for j, i in enumerate(perm):
        assert shuf.tokens[j] == full.tokens[i]
        assert shuf.scores[j] == full.scores[i]
        n = full.frame_lens[i]
        assert np.array_equal(shuf.frame_ids[j, :n], full.frame_ids[i, :n])
    solo = eng.transcribe([waves[4]], return_frames=True)
    assert solo.tokens[0] == full.tokens[4] and solo.scores[0] == full.scores[4]
    again = eng.transcribe(waves)
    assert again.tokens == full.tokens and again.scores == full.scores       # deterministic


def test_pipelined_batches_equal_synchronous_calls(gpu_engines):
    """transcribe_pipelined (double-buffered staging on a copy stream) returns, batch by batch, exactly what transcribe
    returns; shapes change between batches, one batch is empty of frames, the same shape repeats (both slots reused)."""
    eng = gpu_engines()
    rng = np.random.default_rng(5)
    batches = []
    for k in range(7):
        nb = [3, 5, 3, 1, 3, 3, 2][k]
        batches.append([synth.noise_audio(300 + 10 * k + i, int(rng.integers(8000, 60000))) for i in range(nb)])
    batches[3] = [np.zeros(100, np.float32) + 0.01]          # shorter than one frame
    want = [eng.transcribe(b) for b in batches]
    got = list(eng.transcribe_pipelined(iter(batches)))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.tokens == w.tokens and g.scores == w.scores
    assert list(eng.transcribe_pipelined(iter([]))) == []


def test_simt_gemm_path_matches_oracle(gpu_engines):
    """The fp32 FMA-pipe GEMM build of the same layer program (also what the chunk path uses)."""
    from masr_b200.engine import ConformerEngine
    eng = ConformerEngine(synth_weights(0), streaming=True, gemm="simt")
    sd = synth.to_torch(synth_weights(0))
    cfg = oc.ConformerConfig()
    waves = [make_audio("speech", 60, 16000 * 2 + 333), make_audio("noise", 61, 16000)]
    res = eng.transcribe(waves, return_frames=True)
    for i, w in enumerate(waves):
        f = torch.from_numpy(ob.featurize(w.copy()))
        with torch.no_grad():
            probs = oc.get_encoder_out(sd, cfg, f[None])[0].numpy()
        ids, _ = octc.best_path(probs)
        assert np.array_equal(ids, res.frame_ids[i, :res.frame_lens[i]])


def test_edge_cases(gpu_engines):
    eng = gpu_engines()
    # too short for any encoder frame: fewer than 7 feature frames -> empty result, no crash
    res = eng.transcribe([synth.noise_audio(1, 400 + 160 * 5), synth.noise_audio(2, 100)])
    assert res.tokens == [[], []] and res.scores == [0.0, 0.0]
    # mixture of empty-output and normal rows
    res = eng.transcribe([synth.noise_audio(3, 500), synth.speechlike_audio(4, 16000)], return_frames=True)
    assert res.tokens[0] == [] and res.frame_lens.tolist() == [0, 23]


@pytest.fixture(scope="module")
def predictor(tmp_path_factory, predictor_golden):
    import yaml
    from masr_b200.predict import MASRPredictor
    tmp = tmp_path_factory.mktemp("pred")
    g = predictor_golden
    mp = str(tmp / "inference.pt")
    torch.save(synth.to_torch(synth_weights(g["wseed"])), mp)       # plain state_dict checkpoint
    vp, mi = str(tmp / "vocabulary.txt"), str(tmp / "mean_istd.json")
    synth.write_vocabulary(vp)
    synth.write_mean_istd(mi, g["wseed"])
    cfg = {"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy",
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "n_mfcc": 40, "sample_rate": 16000,
                               "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp, "mean_istd_path": mi},
           "ctc_beam_search_decoder_conf": {"beam_size": 300}}
    cfg_path = str(tmp / "conformer.yml")
    with open(cfg_path, "w", encoding="utf-8") as f:
        yaml.safe_dump(cfg, f)
    return MASRPredictor(configs=cfg_path, model_path=mp, use_gpu=True)


def _same(result, want):
    if want is None:
        return result is None
    return result is not None and result["text"] == want["text"] and abs(result["score"] - want["score"]) < SCORE_TOL


def test_predictor_dropin_whole_utterance(predictor, predictor_golden):
    g = predictor_golden
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    assert _same(predictor.predict(audio_data=x.copy()), g["whole"])
    out = predictor.predict_batch([x.copy(), x[:20000].copy()])
    assert _same(out[0], g["whole"])
    outs = list(predictor.predict_batches([[x.copy()], [x[:20000].copy(), x.copy()], [x.copy()]]))
    assert _same(outs[0][0], g["whole"]) and _same(outs[1][1], g["whole"]) and _same(outs[2][0], g["whole"]) and outs[1][0] == out[1]


def test_predictor_dropin_streaming(predictor, predictor_golden):
    g = predictor_golden
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = g["push"]
    for rep in range(2):                                  # twice: reset_stream must restore a clean state
        predictor.reset_stream()
        got = [predictor.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm))
               for s in range(0, len(pcm), push)]
        assert len(got) == len(g["pushes_pcm"])
        for r, w in zip(got, g["pushes_pcm"]):
            assert _same(r, w), (r, w)
    predictor.reset_stream()
    got = [predictor.predict_stream(audio_data=x[s:s + push].copy(), is_end=s + push >= len(x))
           for s in range(0, len(x), push)]
    for r, w in zip(got, g["pushes_ndarray"]):
        assert _same(r, w), (r, w)
    predictor.reset_stream()


def test_predictor_errors(predictor):
    with pytest.raises(Exception):
        predictor.predict(audio_data=12345)
    with pytest.raises(Exception):
        predictor.predict_stream(audio_data="not-bytes")
    with pytest.raises(Exception):
        predictor.predict(audio_data=np.zeros(16000, np.float32), sample_rate=8000)


def test_chunk_path_matches_oracle_states(gpu_engines):
    """forward_chunk parity incl. the caches, chunk by chunk (SURVEY.md §4 'streaming')."""
    eng = gpu_engines()
    sd = synth.to_torch(synth_weights(0))
    cfg = oc.ConformerConfig()
    feat = torch.from_numpy(ob.featurize(make_audio("speech", 8, 16000 * 3 + 4000)))
    fd = feat.to(eng.device)
    st_o = oc.ChunkState()
    st_g = eng.new_stream()
    starts = list(range(0, feat.shape[0] - 67 + 1, 64))
    last = starts[-1] + 64
    chunks = [(s, s + 67) for s in starts] + [(last, feat.shape[0])]      # ragged final chunk (is_end path)
    for (a, b) in chunks:
        if b - a < 7:
            continue
        with torch.no_grad():
            pm = oc.get_encoder_out_chunk(sd, cfg, feat[None, a:b], st_o, -16)[0].numpy()
        ids, maxp, probs = eng.encode_chunk(fd[a:b], st_g, -16, want_probs=True)
        assert np.abs(probs.cpu().numpy() - pm).max() < PROB_TOL
        assert np.array_equal(ids.cpu().numpy(), pm.argmax(1))
        assert st_g.offset == st_o.offset and st_g.cache_len == st_o.att_cache.shape[2]
        # attention cache: oracle [L,h,t,2dk] vs engine [t, k(256)|v(256)]
        kv = st_g.kv[3][st_g.cache_start:st_g.cache_start + st_g.cache_len].cpu()
        ko = st_o.att_cache[3, :, :, :64].permute(1, 0, 2).reshape(-1, 256)
        vo = st_o.att_cache[3, :, :, 64:].permute(1, 0, 2).reshape(-1, 256)
        assert (kv[:, :256] - ko).abs().max().item() < 1e-4 and (kv[:, 256:] - vo).abs().max().item() < 1e-4
        cc = st_g.ws["xcat"][5, :14].cpu()
        assert (cc - st_o.cnn_cache[5, 0].t()).abs().max().item() < 1e-4


def test_ingest_formats_equal_float_input(predictor, tmp_path):
    """a1: every input form `MASRPredictor._load_audio` accepts (predict.py:147-164) — int16 / int32 ndarray,
    stereo ndarray, bytes of a complete WAV file, a path, an open file object — gives exactly the result of the float32
    samples the reference would convert it to (`_convert_samples_to_float32`, audio.py:532-546: integers scaled by
    2^-(bits-1), channels averaged)."""
    import wave
    x = make_audio("speech", 61, 16000 * 2 + 321)
    pcm = (np.clip(x, -1, 1) * 32767).astype(np.int16)
    want = predictor.predict(audio_data=pcm.astype(np.float32) / np.float32(32768.0))
    assert len(want["text"]) > 0
    assert predictor.predict(audio_data=pcm) == want                                     # int16 ndarray
    assert predictor.predict(audio_data=pcm.astype(np.int32) * 65536) == want            # int32 ndarray (same values, 32-bit scale)
    stereo = np.stack([pcm, pcm], axis=1)
    assert predictor.predict(audio_data=stereo) == want                                  # [n, 2] -> channel mean
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    assert predictor.predict(audio_data=path) == want                                    # path
    with open(path, "rb") as f:
        assert predictor.predict(audio_data=f.read()) == want                            # bytes of the whole file
    with open(path, "rb") as f:
        assert predictor.predict(audio_data=f) == want                                   # file object
    with wave.open(path, "wb") as w:                                                     # other sample rate: resampling is out of scope -> raises
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000); w.writeframes(pcm.tobytes())
    with pytest.raises(Exception):
        predictor.predict(audio_data=path)
