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


def run(name, eng, waves, fn, reps=5, oracle=None, sample=(0,)):
    """oracle(wave) -> expected greedy token ids of one utterance on the CPU (the oracle port): the timed call's output is
    checked against it on `sample` before the timing (the config lines carried no correctness check)."""
    if only and name.split()[0] not in only:
        return
    for _ in range(2):
        out = fn(waves)
    verified = None
    if oracle is not None:
        toks = out.tokens if hasattr(out, "tokens") else out[0]
        verified = all(list(toks[i]) == list(oracle(waves[i])) for i in sample)
        assert verified, f"{name}: token ids differ from the CPU oracle"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(waves)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    audio = sum(len(w) for w in waves) / 16000.0
    print(json.dumps({"config": name, "utterances": len(waves), "audio_s": audio, "ms_per_batch": dt * 1e3,
                      "audio_seconds_per_second": audio / dt,
                      "ids_match_cpu_oracle_on_sample": None if verified is None else {"utterances": list(sample), "ok": verified}}), flush=True)


def greedy_oracle(mod, sd, cfg, batched=True):
    from oracle import ctc as octc, fbank as ob

    def f(w):
        with torch.no_grad():
            feat = torch.from_numpy(ob.featurize(w.copy()))
            probs = mod.get_encoder_out(sd, cfg, feat[None] if batched else feat)[0].numpy()
        return octc.collapse(octc.best_path(probs)[0])
    return f


from oracle import conformer as oc, deepspeech2 as ods, efficient_conformer as oe, squeezeformer as osq  # noqa: E402

tens = [synth.noise_audio(1000 + i, 160000) for i in range(32)]
rng = np.random.default_rng(0)
varlen = [synth.noise_audio(2000 + i, int(n)) for i, n in enumerate(rng.integers(16000, 480001, 64))]

sdn = synth.conformer_state_dict(0)
e = ConformerEngine(sdn, streaming=False)
run("config2 conformer.yml streaming=False 32x10s ctc_greedy", e, tens, lambda w: e.transcribe(w),
    oracle=greedy_oracle(oc, synth.to_torch(sdn), oc.ConformerConfig(causal=False)), sample=(0, 31))
del e
sdn = synth.efficient_conformer_state_dict(0)
e = EfficientConformerEngine(sdn, streaming=False)
run("config4 efficient_conformer.yml streaming=False 32x10s/GPU ctc_beam_search(300,40,0.99,no LM)", e, tens, lambda w: e.transcribe_beam(w, **BEAM))
def piped(w, e_=None):
    return list(e.transcribe_beam_pipelined([w] * 4, **BEAM))[-1]


if not only or "config4p" in only:
    # throughput form of config 4: a stream of batches, the beam search of batch k under the encoder of batch k+1
    for _ in range(2):
        piped(tens)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n_b = 12
    res = list(e.transcribe_beam_pipelined([tens] * n_b, **BEAM))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n_b
    same = res[-1][0] == e.transcribe_beam(tens, **BEAM)[0]
    assert same
    print(json.dumps({"config": "config4p efficient_conformer.yml streaming=False 32x10s/GPU ctc_beam_search, PIPELINED stream of batches (beam search of batch k on a second stream under the encoder of batch k+1)",
                      "utterances": 32, "audio_s": 320.0, "ms_per_batch": dt * 1e3, "audio_seconds_per_second": 320.0 / dt,
                      "equals_blocking_call": bool(same)}), flush=True)
run("config4g efficient_conformer.yml streaming=False 32x10s/GPU ctc_greedy", e, tens, lambda w: e.transcribe(w),
    oracle=greedy_oracle(oe, synth.to_torch(sdn), oe.EfficientConfig(causal=False)), sample=(0, 31))
del e
sdn = synth.conformer_state_dict(0)
e = ConformerEngine(sdn, streaming=True)
run("config5 conformer.yml streaming-trained 64 x 1-30s/GPU ctc_beam_search(300,40,0.99,no LM)", e, varlen, lambda w: e.transcribe_beam(w, **BEAM), reps=3)
run("config5g conformer.yml streaming-trained 64 x 1-30s/GPU ctc_greedy", e, varlen, lambda w: e.transcribe(w), reps=3,
    oracle=greedy_oracle(oc, synth.to_torch(sdn), oc.ConformerConfig()), sample=(int(np.argmin([len(w) for w in varlen])), 5))
del e
sdn = synth.squeezeformer_state_dict(0, streaming=True)
e = SqueezeformerEngine(sdn, streaming=True)
run("squeezeformer squeezeformer.yml 32x10s ctc_greedy (whole utterance)", e, tens, lambda w: e.transcribe(w),
    oracle=greedy_oracle(osq, synth.to_torch(sdn), osq.SqueezeformerConfig(causal=True)), sample=(0,))
del e
sdn = synth.deepspeech2_state_dict(0)
e = DeepSpeech2Engine(sdn, streaming=True)
run("deepspeech2 deepspeech2.yml 32x10s ctc_greedy (whole utterance)", e, tens, lambda w: e.transcribe(w),
    oracle=greedy_oracle(ods, synth.to_torch(sdn), ods.DS2Config()), sample=(0,))
