"""Long-form path: the VAD segmentation state machine against the reference's own implementation (frozen by
tests/golden/make_golden.py from masr/infer_utils/vad_predictor.py:106-175 with a scripted probability track in place of
the ONNX network), and ``MASRPredictor.predict_long`` (GPU) against the reference's per-segment loop semantics."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from masr_b200 import vad


def test_segmentation_state_machine_matches_reference_golden():
    with open(os.path.join(GOLDEN, "vad_timestamps_golden.json"), encoding="utf-8") as f:
        cases = json.load(f)
    assert len(cases) >= 10 and any(c["timestamps"] for c in cases) and any(not c["timestamps"] for c in cases)
    for c in cases:
        got = vad.speech_timestamps_from_probs(c["probs"], c["samples"], 16000, **c["kw"])
        assert got == c["timestamps"]
        got2 = vad.ProbabilityVAD(lambda a, sr, p=c["probs"]: p, **c["kw"]).get_speech_timestamps(np.zeros(c["samples"], np.float32), 16000)
        assert got2 == c["timestamps"]


def test_silero_wrapper_fails_loudly_without_onnxruntime():
    try:
        import onnxruntime  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="onnxruntime"):
            vad.SileroVAD("silero_vad.onnx")


@pytest.mark.gpu
def test_predict_long_equals_per_segment_predict(tmp_path):
    """predict.py:195-234: segments -> predict -> '，'.join, mean score rounded to 2 digits; here all segments go through
    one batched pass, which must not change any segment's result."""
    from conftest import make_audio, synth_weights
    from masr_b200 import synth
    from masr_b200.predict import MASRPredictor
    mp, vp = str(tmp_path / "m.pt"), str(tmp_path / "vocabulary.txt")
    torch.save(synth.to_torch(synth_weights(0)), mp)
    synth.write_vocabulary(vp)
    cfg = {"use_model": "conformer", "streaming": True, "decoder": "ctc_greedy",
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp}}
    pred = MASRPredictor(configs=cfg, model_path=mp, use_gpu=True)
    with pytest.raises(Exception, match="VAD"):
        pred.predict_long(np.zeros(16000, np.float32))
    # 30 s recording: three speech-like stretches separated by near-silence; a scripted VAD marks them
    sr, W = 16000, 512
    parts = [make_audio("speech", 500 + i, n) for i, n in enumerate((16000 * 6, 16000 * 9 + 300, 16000 * 4))]
    gap = (np.random.default_rng(0).standard_normal(16000 * 3) * 1e-4).astype(np.float32)
    audio = np.concatenate([gap, parts[0], gap, parts[1], gap, parts[2], gap])
    nwin = (len(audio) + W - 1) // W
    probs = np.full(nwin, 0.05)
    pos = 0
    for seg in (gap, parts[0], gap, parts[1], gap, parts[2], gap):
        if seg is not gap:
            probs[pos // W + 1:(pos + len(seg)) // W] = 0.95
        pos += len(seg)
    v = vad.ProbabilityVAD(lambda a, s: probs)
    stamps = v.get_speech_timestamps(audio, sr)
    assert len(stamps) == 3
    got = pred.predict_long(audio, vad_predictor=v)
    texts, scores = [], []
    for t in stamps:
        r = pred.predict(audio_data=audio[t["start"]:t["end"]].copy())
        if r["text"] != "":
            texts.append(r["text"])
        scores.append(r["score"])
    assert len(texts) >= 2
    assert got == {"text": "，".join(texts), "score": round(sum(scores) / len(scores), 2)}
    # no speech at all: empty text, score 0 (the reference would fail on texts[0]; documented difference)
    assert pred.predict_long(audio, vad_predictor=vad.ProbabilityVAD(lambda a, s: np.zeros(nwin))) == {"text": "", "score": 0}
