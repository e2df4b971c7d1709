"""Dev tool: end-to-end throughput (host float32 waveforms in, token ids + scores out, one blocking call per batch) of the
per-GPU shard of every BASELINE.json config that is not the headline — one JSON line each.  Synthetic audio + weights.
    config 2  conformer.yml streaming=False, 32 x 10 s, ctc_greedy
    config 4  efficient_conformer.yml streaming=False, 32 x 10 s per GPU (256 over 8), ctc_beam_search (no LM)
    config 5  conformer.yml (streaming-trained), 64 utterances of 1-30 s per GPU (512 over 8), ctc_beam_search (no LM)
    plus      squeezeformer.yml / deepspeech2.yml whole-utterance, 32 x 10 s, ctc_greedy
Numbers printed here are dev measurements (CUDA-synchronised wall clock around the public engine call), not bench values."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from masr_b200 import synth
from masr_b200.deepspeech2 import DeepSpeech2Engine
from masr_b200.engine import ConformerEngine, EfficientConformerEngine
from masr_b200.squeezeformer import SqueezeformerEngine

BEAM = dict(beam_size=300, cutoff_prob=0.99, cutoff_top_n=40)
only = set(sys.argv[1:])


def run(name, eng, waves, fn, reps=5):
    if only and name.split()[0] not in only:
        return
    for _ in range(2):
        fn(waves)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(waves)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    audio = sum(len(w) for w in waves) / 16000.0
    print(json.dumps({"config": name, "utterances": len(waves), "audio_s": audio, "ms_per_batch": dt * 1e3,
                      "audio_seconds_per_second": audio / dt}), flush=True)


tens = [synth.noise_audio(1000 + i, 160000) for i in range(32)]
rng = np.random.default_rng(0)
varlen = [synth.noise_audio(2000 + i, int(n)) for i, n in enumerate(rng.integers(16000, 480001, 64))]

e = ConformerEngine(synth.conformer_state_dict(0), streaming=False)
run("config2 conformer.yml streaming=False 32x10s ctc_greedy", e, tens, lambda w: e.transcribe(w))
del e
e = EfficientConformerEngine(synth.efficient_conformer_state_dict(0), streaming=False)
run("config4 efficient_conformer.yml streaming=False 32x10s/GPU ctc_beam_search(300,40,0.99,no LM)", e, tens, lambda w: e.transcribe_beam(w, **BEAM))
run("config4g efficient_conformer.yml streaming=False 32x10s/GPU ctc_greedy", e, tens, lambda w: e.transcribe(w))
del e
e = ConformerEngine(synth.conformer_state_dict(0), streaming=True)
run("config5 conformer.yml streaming-trained 64 x 1-30s/GPU ctc_beam_search(300,40,0.99,no LM)", e, varlen, lambda w: e.transcribe_beam(w, **BEAM), reps=3)
run("config5g conformer.yml streaming-trained 64 x 1-30s/GPU ctc_greedy", e, varlen, lambda w: e.transcribe(w), reps=3)
del e
e = SqueezeformerEngine(synth.squeezeformer_state_dict(0, streaming=True), streaming=True)
run("squeezeformer squeezeformer.yml 32x10s ctc_greedy (whole utterance)", e, tens, lambda w: e.transcribe(w))
del e
e = DeepSpeech2Engine(synth.deepspeech2_state_dict(0), streaming=True)
run("deepspeech2 deepspeech2.yml 32x10s ctc_greedy (whole utterance)", e, tens, lambda w: e.transcribe(w))
