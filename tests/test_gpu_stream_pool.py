"""GPU (-m gpu): many concurrent streams through the batched chunk path == one predict_stream per stream."""
import numpy as np
import pytest
import torch
import yaml

from conftest import make_audio, synth_weights
from masr_b200 import synth

pytestmark = pytest.mark.gpu


def test_stream_pool_equals_single_stream_predictor(tmp_path, predictor_golden):
    from masr_b200.predict import MASRPredictor
    from masr_b200.stream_pool import StreamPool
    mp, vp = str(tmp_path / "m.pt"), str(tmp_path / "vocabulary.txt")
    torch.save(synth.to_torch(synth_weights(0)), mp)
    synth.write_vocabulary(vp)
    cfg = {"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy",
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp}}
    pred = MASRPredictor(configs=cfg, model_path=mp, use_gpu=True)
    g = predictor_golden
    audios = [make_audio(g["kind"], g["aseed"], g["samples"]),                 # the stream frozen from the reference
              make_audio("noise", 91, 16000 * 2 + 4000), make_audio("speech", 92, 16000 * 3 + 1234)]
    pcms = [(np.clip(a, -1, 1) * 32767).astype("<i2") for a in audios]
    push = 8000
    # reference behaviour: one predictor, one stream at a time
    want = []
    for pcm in pcms:
        pred.reset_stream()
        want.append([pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)])
    pred.reset_stream()
    for r, w in zip(want[0], g["pushes_pcm"]):                                 # sanity: stream 0 still equals the reference golden
        assert (r is None) == (w is None) and (r is None or r["text"] == w["text"])
    # all three streams concurrently
    pool = StreamPool(pred.predictor, synth.vocabulary(), n_slots=4)           # one spare slot stays idle
    got = [[] for _ in pcms]
    npush = [len(range(0, len(p), push)) for p in pcms]
    for k in range(max(npush)):
        mid = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(len(pcms)) if k < npush[i] - 1}
        last = {i: pcms[i][k * push:(k + 1) * push].tobytes() for i in range(len(pcms)) if k == npush[i] - 1}
        if mid:
            out = pool.push(mid, is_end=False)
            for i in mid:
                got[i].append(out[i])
        if last:
            out = pool.push(last, is_end=True)
            for i in last:
                got[i].append(out[i])
    for i in range(len(pcms)):
        assert len(got[i]) == len(want[i])
        for r, w in zip(got[i], want[i]):
            assert (r is None) == (w is None), (i, r, w)
            if r is not None:
                assert r["text"] == w["text"], (i, r, w)
                assert abs(r["score"] - w["score"]) < 1e-3
    # slot reuse after reset
    pool.reset_stream(1)
    out = pool.push({1: pcms[1][:push * 3].tobytes()}, is_end=True)
    pred.reset_stream()
    ref = pred.predict_stream(audio_data=pcms[1][:push * 3].tobytes(), is_end=True)
    assert out[1]["text"] == ref["text"]
