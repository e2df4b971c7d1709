import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import make_audio, synth_weights
from masr_b200 import synth
from masr_b200.engine import ConformerEngine
from oracle import beam as obeam, fbank as ob
from test_gpu_configs import ctc_loglik, BEAM
eng = ConformerEngine(synth_weights(0), streaming=True)
w = make_audio("speech", 132, 203117)
toks, scores = eng.transcribe_beam([w], **BEAM)
feats, frames, status = eng.fbank([w])
toks2, scores2 = eng.beam_features(feats, frames, **BEAM)
print("same via features:", toks2[0] == toks[0], scores, scores2)
probs = eng.posteriors(feats.cpu().numpy(), frames)[0]
(score, want), = obeam.prefix_beam_search(probs, **BEAM)
print("gpu len", len(toks[0]), "cpu len", len(want), "equal", toks[0] == want, "gpu score", scores[0], "cpu score", score)
print("loglik gpu", ctc_loglik(probs, toks[0]), "cpu", ctc_loglik(probs, want))
nb = obeam.prefix_beam_search(probs, nbest=5, **BEAM)
for sc, tk in nb:
    print("cpu nbest", sc, len(tk), tk == toks[0])
d = [i for i, (a, b) in enumerate(zip(toks[0], want)) if a != b]
print("first diffs", d[:5], toks[0][80:92], want[80:92])
