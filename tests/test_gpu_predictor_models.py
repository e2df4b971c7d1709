"""GPU (-m gpu): the MASRPredictor drop-in over every model family and both decoders."""
import numpy as np
import pytest
import torch
import yaml

from conftest import make_audio
from masr_b200 import synth
from oracle import beam as obeam, conformer as oc, ctc as octc, deepspeech2 as od, efficient_conformer as oe, fbank as ob, squeezeformer as osq

pytestmark = pytest.mark.gpu


def build(tmp, use_model, sd, streaming=True, decoder="ctc_greedy"):
    from masr_b200.predict import MASRPredictor
    mp, vp = str(tmp / f"{use_model}.pt"), str(tmp / "vocabulary.txt")
    torch.save(synth.to_torch(sd), mp)
    synth.write_vocabulary(vp)
    cfg = {"use_model": use_model, "streaming": streaming, "decoder": decoder,
           "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
           "dataset_conf": {"dataset_vocab": vp},
           "ctc_beam_search_decoder_conf": {"alpha": 2.2, "beta": 4.3, "beam_size": 20, "cutoff_prob": 0.99, "cutoff_top_n": 40,
                                            "num_processes": 10, "language_model_path": "lm/none.klm"}}
    p = str(tmp / f"{use_model}.yml")
    with open(p, "w", encoding="utf-8") as f:
        yaml.safe_dump(cfg, f)
    return MASRPredictor(configs=p, model_path=mp, use_gpu=True)


@pytest.mark.parametrize("use_model", ["efficient_conformer", "squeezeformer", "deepspeech2"])
def test_predict_matches_oracle(tmp_path, use_model):
    vocab = synth.vocabulary()
    if use_model == "efficient_conformer":
        sd = synth.efficient_conformer_state_dict(0)
        ref = lambda f: oe.get_encoder_out(synth.to_torch(sd), oe.EfficientConfig(), f)[0]
    elif use_model == "squeezeformer":
        sd = synth.squeezeformer_state_dict(0)
        ref = lambda f: osq.get_encoder_out(synth.to_torch(sd), osq.SqueezeformerConfig(), f)[0]
    else:
        sd = synth.deepspeech2_state_dict(0)
        ref = lambda f: od.get_encoder_out(synth.to_torch(sd), od.DS2Config(), f)[0]
    pred = build(tmp_path, use_model, sd)
    x = make_audio("speech", 77, 16000 * 2 + 500)
    with torch.no_grad():
        probs = ref(torch.from_numpy(ob.featurize(x.copy()))[None]).numpy()
    score, text, _ = octc.greedy_decode(probs, vocab)
    out = pred.predict(audio_data=x.copy())
    assert out["text"] == text and abs(out["score"] - score) < 1e-3
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    if use_model == "deepspeech2":
        pred.reset_stream()
        got = [pred.predict_stream(audio_data=pcm[s:s + 8000].tobytes(), is_end=s + 8000 >= len(pcm)) for s in range(0, len(pcm), 8000)]
        assert got[-1] is not None and isinstance(got[-1]["text"], str)
    # (the Squeezeformer / EfficientConformer streaming paths are pinned to the reference in
    #  tests/test_squeezeformer_stream.py and tests/test_efficient_stream.py)


def test_beam_search_decoder_config(tmp_path):
    sd = synth.conformer_state_dict(0)
    pred = build(tmp_path, "conformer", sd, decoder="ctc_beam_search")
    x = make_audio("speech", 78, 16000 * 2)
    with torch.no_grad():
        probs = oc.get_encoder_out(synth.to_torch(sd), oc.ConformerConfig(), torch.from_numpy(ob.featurize(x.copy()))[None])[0].numpy()
    (score, toks), = obeam.prefix_beam_search(probs, beam_size=20, cutoff_prob=0.99, cutoff_top_n=40)
    out = pred.predict(audio_data=x.copy())
    assert out["text"] == octc.ids_to_text(toks, synth.vocabulary())
    assert abs(out["score"] - score) < 5e-3 * max(1.0, abs(score))
    outs = pred.predict_batch([x.copy(), x[:20000].copy()])
    assert outs[0]["text"] == out["text"]


def test_deepspeech2_predict_stream_matches_reference_golden(tmp_path):
    """The streaming (unidirectional) DeepSpeech2 through ``MASRPredictor.predict_stream`` against the pushes frozen from the
    reference's real predictor (tests/golden/predictor_golden_deepspeech2.json)."""
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "predictor_golden_deepspeech2.json"), encoding="utf-8") as f:
        g = json.load(f)
    pred = build(tmp_path, "deepspeech2", synth.deepspeech2_state_dict(g["wseed"], streaming=True))
    x = make_audio(g["kind"], g["aseed"], g["samples"])
    whole = pred.predict(audio_data=x.copy())
    assert whole["text"] == g["whole"]["text"] and abs(whole["score"] - g["whole"]["score"]) < 1e-3
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = g["push"]
    pred.reset_stream()
    got = [pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)]
    assert len(got) == len(g["pushes_pcm"])
    for r, w in zip(got, g["pushes_pcm"]):
        assert (r is None) == (w is None), (r, w)
        if r is not None:
            assert r["text"] == w["text"], (r, w)
            assert abs(r["score"] - w["score"]) < 1e-3
