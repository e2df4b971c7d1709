"""GPU (-m gpu): scaled-down analogues of BASELINE.json's configs 4 and 5 — the same code path a sharded 8-GPU run takes on
each rank (partition -> engine.transcribe_beam on the shard -> ordered results), checked against the oracle:
  config 4  efficient_conformer.yml non-streaming, beam search (beam 300, top-n 40, cutoff 0.99, no LM), ragged batch
  config 5  conformer.yml (streaming-trained, causal), variable length 1-30 s, beam search
The beam search itself is parity-unpinned against the reference (external decoder absent); the GPU path must
equal the CPU restatement run on the oracle's posteriors."""
import numpy as np
import pytest
import torch

from conftest import make_audio, synth_weights
from masr_b200 import shard, synth
from oracle import beam as obeam, conformer as oc, efficient_conformer as oe, fbank as ob

pytestmark = pytest.mark.gpu
BEAM = dict(beam_size=300, cutoff_prob=0.99, cutoff_top_n=40)


def ctc_loglik(probs, toks, blank=0):
    """log P(tokens | posteriors): the CTC forward algorithm in float64 (no pruning)."""
    ext = [blank]
    for t in toks:
        ext += [int(t), blank]
    L = len(ext)
    lp = np.log(np.maximum(probs.astype(np.float64), 1e-300))
    a = np.full(L, -np.inf)
    a[0] = lp[0, blank]
    if L > 1:
        a[1] = lp[0, ext[1]]
    for t in range(1, probs.shape[0]):
        b = a.copy()
        b[1:] = np.logaddexp(b[1:], a[:-1])
        for s_ in range(2, L):
            if ext[s_] != blank and ext[s_] != ext[s_ - 2]:
                b[s_] = np.logaddexp(b[s_], a[s_ - 2])
        a = b + lp[t, ext]
    return float(np.logaddexp(a[-1], a[-2]) if L > 1 else a[-1])


def check(eng, waves, ref_probs):
    """Encoder parity against the oracle (per-frame argmax exact, posteriors within 5e-5), then the GPU beam search against
    the CPU restatement run on the SAME posteriors (the engine's): with near-tied hypotheses over hundreds of frames a 1e-6
    difference in the inputs legitimately changes a pruned beam search, so decoder and encoder are compared separately."""
    toks, scores = shard.sharded_transcribe(waves, lambda ws: eng.transcribe_beam(ws, **BEAM), max_tokens=800)
    assert len(toks) == len(waves)
    order = shard.partition([len(w) for w in waves], 1)[0]          # one rank: the shard is the whole list, longest first
    cands = dict(zip(order, eng.last_beam_candidates()))
    for i, w in enumerate(waves):
        feat = ob.featurize(w.copy())
        want_probs = ref_probs(torch.from_numpy(feat)[None])
        probs = eng.posteriors(feat[None], [feat.shape[0]])[0]
        assert probs.shape == want_probs.shape
        assert np.array_equal(probs.argmax(1), want_probs.argmax(1)), i
        assert np.abs(probs - want_probs).max() < 5e-5, i
        # the search, bit for bit: restatement on the candidate lists the GPU searched over (same log-sum-exp operation
        # sequence on both sides, oracle/beam.py) -> identical prefix, identical float32 score, at any utterance length
        (score, want), = obeam.prefix_beam_search(probs, cands_per_frame=cands[i][:probs.shape[0]], **BEAM)
        assert toks[i] == want, i
        assert np.float32(scores[i]) == np.float32(score), i
        # and the pruning: candidate ids per frame equal the restatement's prune of the engine's posteriors
        for t in (0, probs.shape[0] // 2, probs.shape[0] - 1):
            assert [c for c, _ in cands[i][t]] == [c for c, _ in obeam.prune_frame(probs[t], BEAM["cutoff_prob"], BEAM["cutoff_top_n"])], (i, t)


def test_config4_efficient_conformer_nonstreaming_beam():
    from masr_b200.engine import EfficientConformerEngine
    sdn = synth.efficient_conformer_state_dict(1)
    eng = EfficientConformerEngine(sdn, streaming=False)
    sd = synth.to_torch(sdn)
    cfg = oe.EfficientConfig(causal=False)
    lens = [16000 * 3 + 5, 16000 * 2, 9000, 16000 * 4 + 321, 16000 + 160]
    waves = [make_audio("speech" if i % 2 else "noise", 120 + i, n) for i, n in enumerate(lens)]

    def ref(feat):
        with torch.no_grad():
            return oe.get_encoder_out(sd, cfg, feat)[0].numpy()
    check(eng, waves, ref)


def test_config5_conformer_variable_length_1_to_30s_beam(gpu_engines):
    eng = gpu_engines(0, True)
    sd = synth.to_torch(synth_weights(0))
    cfg = oc.ConformerConfig()
    lens = [16000, 480000, 203117]          # the shortest and longest utterance of the config and one in between
    waves = [make_audio("noise" if i % 2 else "speech", 130 + i, n) for i, n in enumerate(lens)]

    def ref(feat):
        with torch.no_grad():
            return oc.get_encoder_out(sd, cfg, feat)[0].numpy()
    check(eng, waves, ref)


def test_pipelined_beam_equals_blocking_calls(gpu_engines):
    """`transcribe_beam_pipelined`: the prefix beam search of batch k on a second stream under the encoder of batch k+1 (two
    buffer sets, results one batch late) returns exactly what one blocking `transcribe_beam` per batch returns — ragged batches,
    changing batch sizes, an empty batch and a batch of too-short audio included."""
    eng = gpu_engines(0, True)
    rng = np.random.default_rng(11)
    batches = []
    for k, B in enumerate((5, 3, 0, 7, 2, 6)):
        batches.append([make_audio("speech" if (i + k) % 2 else "noise", 900 + 10 * k + i, int(rng.integers(8000, 16000 * 4))) for i in range(B)])
    batches.append([make_audio("noise", 990, 300)])                      # shorter than one frame: no output frames at all
    want = [eng.transcribe_beam(b, **BEAM) if b else ([], []) for b in batches]
    got = list(eng.transcribe_beam_pipelined(iter(batches), **BEAM))
    assert len(got) == len(want)
    for (gt, gs), (wt, ws) in zip(got, want):
        assert gt == wt
        assert [np.float32(x) for x in gs] == [np.float32(x) for x in ws]
    assert sum(len(t) for gt, _ in got for t in gt) > 10
