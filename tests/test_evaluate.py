"""Batched evaluation caller: error-rate metrics against hand-computed values (CPU) and the manifest loop
through the CUDA path (GPU)."""
import json

import numpy as np
import pytest
import torch

from masr_b200 import evaluate as ev


def test_levenshtein_and_rates():
    assert ev.levenshtein("kitten", "sitting") == 3
    assert ev.levenshtein("", "abc") == 3 and ev.levenshtein("abc", "abc") == 0
    assert ev.cer("今 天 天气", "今天天气好") == pytest.approx(1 / 5)
    assert ev.cer("abc", "abc") == 0.0
    assert ev.wer("the cat sat", "the cat sat down") == pytest.approx(1 / 4)
    assert ev.wer("a b c", "a x c") == pytest.approx(1 / 3)
    vocab = ["<blank>", "<unk>", "a", "b", "<space>", "<eos>"]
    assert ev.labels_to_string([[2, 4, 3, 5, -1, -1], [3, 3, -1, -1, -1, -1]], vocab, eos=5) == ["a b", "bb"]


def test_read_manifest(tmp_path):
    p = tmp_path / "manifest.test"
    p.write_text(json.dumps({"audio_filepath": "a.wav", "text": "你好", "duration": 1.0}, ensure_ascii=False) + "\n\n" +
                 json.dumps({"audio_filepath": "b.wav", "text": "x y"}) + "\n", encoding="utf-8")
    assert list(ev.read_manifest(str(p))) == [("a.wav", "你好"), ("b.wav", "x y")]


@pytest.mark.gpu
def test_evaluate_loop_on_gpu(tmp_path):
    """Labels = the CPU ORACLE's transcripts (fbank -> encoder -> greedy, the reference's B=1 loop) with one utterance
    perturbed: the mean CER of the CUDA path over the manifest loop is exactly that perturbation's share, i.e. every GPU
    transcript equals the oracle's (the earlier version compared the path with itself)."""
    import yaml
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
    audios = [make_audio("speech", 300 + i, 16000 * 2 + 777 * i) for i in range(7)]
    from oracle import conformer as oc, ctc as octc, fbank as ob
    sd, vocab = synth.to_torch(synth_weights(0)), synth.vocabulary()
    texts = []
    with torch.no_grad():
        for a in audios:
            probs = oc.get_encoder_out(sd, oc.ConformerConfig(), torch.from_numpy(ob.featurize(a.copy()))[None])[0].numpy()
            texts.append(octc.greedy_decode(probs, vocab)[1])
    assert all(len(t) > 1 for t in texts)
    assert [pred.predict(audio_data=a.copy())["text"] for a in audios] == texts      # the single-utterance API agrees too
    labels = list(texts)
    labels[3] = labels[3][1:]                       # drop one character of one reference
    err, n = ev.evaluate(pred, zip(audios, labels), batch_size=3, metrics_type="cer")
    assert n == 7
    assert err == pytest.approx((1.0 / len(labels[3].replace(" ", ""))) / 7)
