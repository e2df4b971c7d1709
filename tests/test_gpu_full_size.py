"""GPU (-m gpu): every BASELINE.json config pinned AT ITS STATED SIZE (VERDICT r1 "What's weak" 1).

  config 1 (headline)  conformer.yml streaming (causal), 32 x 10 s, ctc_greedy: all 32 utterances' per-frame ids bit-exact
                       against the oracle (the reference's own B=1 loop), through the CUDA-graph step bench.py times AND
                       through the pipelined public API (`MASRPredictor.predict_batches`);
  config 2             conformer.yml non-streaming (non-causal), 32 x 10 s: same bar;
  config 3             squeezeformer.yml streaming, 64 live streams x 0.5 s pushes: every stream of the 64-slot pool returns,
                       push by push, what one `predict_stream` per stream returns; a sample of streams (incl. the utterance
                       frozen from the reference) is checked against the oracle / the reference golden;
  configs 4 / 5        per-GPU shard sizes (32 x 10 s EfficientConformer non-streaming; 64 utterances of 1-30 s causal
                       Conformer): per-frame argmax of the posteriors the beam search consumes bit-exact against the oracle on
                       a sample of the shard, greedy ids of the whole shard independent of the batch composition, and the GPU
                       prefix beam search equal to the CPU restatement on the engine's own posteriors (beam: parity unpinned).
The oracle legs are sized for a few minutes of CPU in total.
"""
import numpy as np
import pytest
import torch

from conftest import make_audio, synth_weights
from masr_b200 import synth
from oracle import conformer as oc, ctc as octc, fbank as ob

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-3


def bench_waves(rank=0, n=32):
    """The batch bench.py times (bench.make_waves): 32 x 10 s of seeded noise per rank."""
    return [synth.noise_audio(1000 * rank + i, 160000) for i in range(n)]


def oracle_pass(sd, cfg, waves, vocab):
    out = []
    with torch.no_grad():
        for w in waves:
            feat = torch.from_numpy(ob.featurize(w.copy()))
            probs = oc.get_encoder_out(sd, cfg, feat[None])[0].numpy()
            ids, _ = octc.best_path(probs)
            score, _, toks = octc.greedy_decode(probs, vocab)
            out.append((ids, toks, score))
    return out


@pytest.mark.parametrize("streaming,wseed", [(True, 0), (False, 1)], ids=["config1_headline_causal", "config2_noncausal"])
def test_32x10s_full_batch_ids_bit_exact(gpu_engines, streaming, wseed, tmp_path):
    eng = gpu_engines(wseed, streaming)
    # config 1: exactly the batch bench.py times; config 2: every other utterance speech-like (the non-causal synthetic model
    # answers pure noise with blanks only, which would make the token comparison vacuous)
    waves = bench_waves() if streaming else [make_audio("speech" if i % 2 else "noise", 1000 + i, 160000) for i in range(32)]
    res = eng.transcribe(waves, return_frames=True)              # the CUDA-graph device step of bench.py's `value`
    ref = oracle_pass(synth.to_torch(synth_weights(wseed)), oc.ConformerConfig(causal=streaming), waves, synth.vocabulary())
    mism = 0
    for i, (ids, toks, score) in enumerate(ref):
        n = int(res.frame_lens[i])
        assert n == len(ids) == 248
        mism += int((ids != res.frame_ids[i, :n]).sum())
        assert toks == res.tokens[i], i
        assert abs(score - res.scores[i]) < SCORE_TOL, i
    assert mism == 0                                             # 32 x 248 = 7936 frames, all bit-exact
    assert sum(len(t) for t in res.tokens) >= 16                 # ... and the comparison is not vacuous (non-blank output)
    # the same batch through the user-facing pipelined API (bench.py's `e2e`)
    from masr_b200.predict import MASRPredictor
    mp, vp = str(tmp_path / "m.pt"), str(tmp_path / "vocabulary.txt")
    torch.save(synth.to_torch(synth_weights(wseed)), mp)
    synth.write_vocabulary(vp)
    cfg = {"use_model": "conformer", "streaming": streaming, "decoder": "ctc_greedy",
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp}}
    pred = MASRPredictor(configs=cfg, model_path=mp, use_gpu=True)
    vocab = synth.vocabulary()
    outs = list(pred.predict_batches([waves, waves[::-1]]))
    for got, order in zip(outs, (range(32), range(31, -1, -1))):
        for j, i in enumerate(order):
            assert got[j]["text"] == octc.ids_to_text(ref[i][1], vocab)
            assert abs(got[j]["score"] - ref[i][2]) < SCORE_TOL


def test_config3_squeezeformer_64_live_streams(tmp_path):
    import json
    import os
    from conftest import GOLDEN
    from masr_b200.stream_pool import StreamPool
    from oracle import squeezeformer as osq
    from test_squeezeformer_stream import build_predictor, oracle_predict_stream
    with open(os.path.join(GOLDEN, "predictor_golden_squeezeformer.json"), encoding="utf-8") as f:
        golden = json.load(f)
    weights = synth.squeezeformer_state_dict(golden["wseed"], streaming=True)
    pred = build_predictor(tmp_path, weights)
    S, push = 64, 8000                                            # 64 live streams, 0.5 s pushes
    rng = np.random.default_rng(3)
    lens = [int(n) for n in rng.integers(16000 * 2, 16000 * 5, S)]
    audios = [make_audio("speech" if s % 2 else "noise", 300 + s, lens[s]) for s in range(S)]
    audios[0] = make_audio(golden["kind"], golden["aseed"], golden["samples"])        # the stream frozen from the reference
    pcms = [(np.clip(a, -1, 1) * 32767).astype("<i2") for a in audios]
    assert golden["push"] == push
    # reference behaviour: one predictor, one stream at a time
    want = []
    for pcm in pcms:
        pred.reset_stream()
        want.append([pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)])
    pred.reset_stream()
    for r, w in zip(want[0], golden["pushes_pcm"]):               # stream 0 == the reference's frozen pushes
        assert (r is None) == (w is None) and (r is None or (r["text"] == w["text"] and abs(r["score"] - w["score"]) < 1e-3))
    sd, cfg, vocab = synth.to_torch(weights), osq.SqueezeformerConfig(causal=True), synth.vocabulary()
    for i in (1, 17, 40, 63):                                     # a sample of streams against the CPU oracle
        ref = oracle_predict_stream(sd, cfg, pcms[i], push, vocab)
        for r, w in zip(want[i], ref):
            assert (r is None) == (w is None) and (r is None or (r["text"] == w["text"] and abs(r["score"] - w["score"]) < 1e-3)), i
    # all 64 streams concurrently through the pool
    pool = StreamPool(pred.predictor, vocab, n_slots=S, max_frames=max(len(p) for p in pcms) // 640 + 64)
    got = [[] for _ in pcms]
    npush = [len(range(0, len(p), push)) for p in pcms]
    for k in range(max(npush)):
        mid = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(S) if k < npush[i] - 1}
        last = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(S) if k == npush[i] - 1}
        for grp, is_end in ((mid, False), (last, True)):
            if grp:
                out = pool.push(grp, is_end=is_end)
                for i in grp:
                    got[i].append(out[i])
    nonempty = 0
    for i in range(S):
        assert len(got[i]) == len(want[i])
        for r, w in zip(got[i], want[i]):
            assert (r is None) == (w is None), (i, r, w)
            if r is not None:
                assert r["text"] == w["text"], (i, r, w)
                assert abs(r["score"] - w["score"]) < 1e-3
                nonempty += len(r["text"]) > 0
    assert nonempty > S                                            # the comparison is not vacuous


def _beam_equals_restatement(eng, waves, sample):
    from oracle import beam as obeam
    from test_gpu_configs import BEAM
    toks, scores = eng.transcribe_beam(waves, **BEAM)
    cands = eng.last_beam_candidates()
    for i in sample:
        feat = ob.featurize(waves[i].copy())
        probs = eng.posteriors(feat[None], [feat.shape[0]])[0]
        T = probs.shape[0]                           # frames of utterance i (rows beyond it in the padded batch are ignored)
        # bit for bit on the candidates the GPU searched over (see tests/test_gpu_configs.py::check)
        (score, want), = obeam.prefix_beam_search(probs, cands_per_frame=cands[i][:T], **BEAM)
        assert toks[i] == want, i
        assert np.float32(scores[i]) == np.float32(score), i


def test_config4_shard_efficient_conformer_32x10s_nonstreaming():
    """One rank's shard of config 4 (256 x 10 s over 8 GPUs = 32 x 10 s per GPU)."""
    from masr_b200.engine import EfficientConformerEngine
    from oracle import efficient_conformer as oe
    sdn = synth.efficient_conformer_state_dict(1)
    eng = EfficientConformerEngine(sdn, streaming=False)
    sd, cfg = synth.to_torch(sdn), oe.EfficientConfig(causal=False)
    waves = bench_waves(rank=3)
    full = eng.transcribe(waves, return_frames=True)
    sample = (0, 13, 31)
    for i in sample:
        feat = ob.featurize(waves[i].copy())
        with torch.no_grad():
            want = oe.get_encoder_out(sd, cfg, torch.from_numpy(feat)[None])[0].numpy()
        n = int(full.frame_lens[i])
        assert n == want.shape[0] == 124
        assert np.array_equal(full.frame_ids[i, :n], want.argmax(1)), i
        probs = eng.posteriors(feat[None], [feat.shape[0]])[0]
        assert np.abs(probs - want).max() < 5e-5
    half = eng.transcribe(waves[16:] + waves[:16], return_frames=True)            # batch composition must not matter
    for j in range(32):
        i = (j + 16) % 32
        assert half.tokens[j] == full.tokens[i] and np.array_equal(half.frame_ids[j, :124], full.frame_ids[i, :124])
    _beam_equals_restatement(eng, waves, sample)


def test_config5_shard_conformer_64_utterances_1_to_30s(gpu_engines):
    """One rank's shard of config 5 (512 utterances of 1-30 s over 8 GPUs = 64 per GPU), causal Conformer."""
    eng = gpu_engines(0, True)
    rng = np.random.default_rng(5)
    lens = [16000, 480000] + [int(n) for n in rng.integers(16000, 480001, 62)]
    speech = [i % 3 == 0 or i == 1 for i in range(64)]
    waves = [make_audio("speech" if speech[i] else "noise", 700 + i, n) for i, n in enumerate(lens)]
    full = eng.transcribe(waves, return_frames=True)
    sd, cfg, vocab = synth.to_torch(synth_weights(0)), oc.ConformerConfig(), synth.vocabulary()
    by_len = [int(i) for i in np.argsort(lens)]
    mid = next(i for i in by_len[28:] if speech[i])                # a speech-like utterance of about the median length
    sample = [0, 1, mid, 63]                                       # shortest, longest, ~median (all speech-like), the last (noise)
    ref = oracle_pass(sd, cfg, [waves[i] for i in sample], vocab)
    for i, (ids, toks, score) in zip(sample, ref):
        n = int(full.frame_lens[i])
        assert n == len(ids)
        assert np.array_equal(ids, full.frame_ids[i, :n]), i
        assert toks == full.tokens[i] and abs(score - full.scores[i]) < SCORE_TOL
    perm = list(rng.permutation(64))
    shuf = eng.transcribe([waves[i] for i in perm], return_frames=True)
    for j, i in enumerate(perm):
        n = int(full.frame_lens[i])
        assert shuf.tokens[j] == full.tokens[i] and np.array_equal(shuf.frame_ids[j, :n], full.frame_ids[i, :n])
    _beam_equals_restatement(eng, waves, sample)
