"""CTC prefix beam search (no LM; parity unpinned — the reference's external decoder is absent): the CPU restatement is
checked against its defining properties, and the CUDA kernels against the restatement."""
import numpy as np
import pytest
import torch

from oracle import beam as obeam, ctc as octc


def rand_posteriors(seed, T, V, peaky=6.0, blank_boost=2.0):
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((T, V)).astype(np.float32) * peaky
    logits[:, 0] += blank_boost
    # temporal smoothness so repeats occur
    for t in range(1, T):
        logits[t] = 0.5 * logits[t] + 0.5 * logits[t - 1]
    e = np.exp(logits - logits.max(1, keepdims=True))
    return (e / e.sum(1, keepdims=True)).astype(np.float32), logits


def test_pruning_rule():
    p = np.array([0.05, 0.5, 0.3, 0.1, 0.05], np.float32)
    assert [c for c, _ in obeam.prune_frame(p, 0.99, 40)] == [1, 2, 3, 0, 4]
    assert [c for c, _ in obeam.prune_frame(p, 0.85, 40)] == [1, 2, 3]
    assert [c for c, _ in obeam.prune_frame(p, 0.99, 2)] == [1, 2]
    assert [c for c, _ in obeam.prune_frame(np.array([0.25, 0.25, 0.25, 0.25], np.float32), 0.5, 40)] == [0, 1]   # ties: lower id first


def test_beam1_on_one_hot_equals_greedy():
    rng = np.random.default_rng(0)
    T, V = 40, 50
    ids = rng.integers(0, 6, T)
    probs = np.full((T, V), 1e-9, np.float32)
    probs[np.arange(T), ids] = 1.0
    (score, toks), = obeam.prefix_beam_search(probs, beam_size=1, cutoff_prob=1.0, cutoff_top_n=V)
    assert toks == octc.collapse(ids)
    assert abs(score) < 1e-4


def test_wider_beam_never_scores_worse_and_sums_paths():
    probs, _ = rand_posteriors(1, 25, 30)
    s1 = obeam.prefix_beam_search(probs, beam_size=1, cutoff_prob=1.0, cutoff_top_n=30)[0][0]
    s8 = obeam.prefix_beam_search(probs, beam_size=8, cutoff_prob=1.0, cutoff_top_n=30)[0][0]
    s64 = obeam.prefix_beam_search(probs, beam_size=64, cutoff_prob=1.0, cutoff_top_n=30)[0][0]
    assert s8 >= s1 - 1e-5 and s64 >= s8 - 1e-5
    # exact check on a tiny case: the prefix probability is the sum over all alignments that collapse to it
    T, V = 4, 3
    p, _ = rand_posteriors(2, T, V, peaky=1.0, blank_boost=0.0)
    import itertools
    mass = {}
    for path in itertools.product(range(V), repeat=T):
        key = tuple(octc.collapse(path))
        mass[key] = mass.get(key, 0.0) + float(np.prod([p[t, c] for t, c in enumerate(path)]))
    best = max(mass.items(), key=lambda kv: kv[1])
    (score, toks), = obeam.prefix_beam_search(p, beam_size=50, cutoff_prob=1.0, cutoff_top_n=V)
    assert tuple(toks) == best[0] and abs(np.exp(score) - best[1]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("seed,T,beam,topn,cut", [(3, 60, 300, 40, 0.99), (4, 37, 16, 40, 0.99), (5, 80, 300, 10, 1.0), (6, 25, 1, 40, 0.99)])
def test_gpu_beam_equals_restatement(seed, T, beam, topn, cut):
    from masr_b200 import _lib
    _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    V = 4233
    probs, logits = rand_posteriors(seed, T, V)
    lens = [T, max(1, T // 2)]
    B = len(lens)
    ldl = (V + 15) // 16 * 16
    L = torch.zeros(B * T, ldl, device=dev)
    L[:T, :V] = torch.from_numpy(logits).to(dev)
    L[T:2 * T, :V] = torch.from_numpy(logits[::-1].copy()).to(dev)          # second utterance: the reversed sequence
    st = torch.cuda.current_stream().cuda_stream
    cid = torch.empty(B * T, 40, dtype=torch.int32, device=dev); clp = torch.empty(B * T, 40, device=dev)
    cn = torch.empty(B * T, dtype=torch.int32, device=dev)
    _lib.call("masr_ctc_topk_f32", L.data_ptr(), ldl, B * T, V, topn, cut, cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), st)
    # candidate lists equal the restatement's pruning
    ph = torch.softmax(L[:, :V], 1).cpu().numpy()
    cid_h, cn_h = cid.cpu().numpy(), cn.cpu().numpy()
    for r in (0, T // 2, T - 1, T + 3):
        want = [c for c, _ in obeam.prune_frame(ph[r], cut, topn)]
        assert cid_h[r, :cn_h[r]].tolist() == want
    import ctypes as C
    pool_n, trie_n = C.c_int64(0), C.c_int64(0)
    _lib.call("masr_ctc_prefix_beam_workspace", B, T, C.byref(pool_n), C.byref(trie_n))
    pool = torch.empty(pool_n.value, device=dev); tp = torch.empty(B * trie_n.value, dtype=torch.int32, device=dev)
    tt = torch.empty_like(tp)
    ld = torch.tensor(lens, dtype=torch.int32, device=dev)
    otok = torch.zeros(B, T, dtype=torch.int32, device=dev); on = torch.zeros(B, dtype=torch.int32, device=dev)
    osc = torch.zeros(B, device=dev)
    _lib.call("masr_ctc_prefix_beam", cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), T, ld.data_ptr(), B, beam, 0, pool.data_ptr(),
              tp.data_ptr(), tt.data_ptr(), trie_n.value, otok.data_ptr(), T, on.data_ptr(), osc.data_ptr(), st)
    torch.cuda.synchronize()
    clp_h = clp.cpu().numpy()
    for b in range(B):
        p = ph[b * T: b * T + lens[b]]
        # the SEARCH, bit for bit: the restatement on the kernel's own candidate lists (both evaluate log-sum-exp in the same
        # specified sequence of float32 operations) -> identical prefix and identical float32 score
        cands = [[(int(cid_h[b * T + t, k]), clp_h[b * T + t, k]) for k in range(cn_h[b * T + t])] for t in range(lens[b])]
        (score, toks), = obeam.prefix_beam_search(p, beam_size=beam, cutoff_prob=cut, cutoff_top_n=topn, cands_per_frame=cands)
        got = otok[b, :on[b].item()].cpu().tolist()
        assert got == toks, (b, got, toks)
        assert np.float32(osc[b].item()) == np.float32(score), (b, osc[b].item(), score)
        # the candidate log-probabilities themselves: within float32 rounding of log(softmax)
        for t in (0, lens[b] - 1):
            for c, lp in cands[t]:
                assert abs(float(lp) - np.log(float(p[t, c]))) < 2e-6 * max(1.0, abs(float(lp)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,T,beam,chunk", [(7, 75, 300, 16), (8, 50, 32, 7), (9, 33, 300, 33)])
def test_gpu_streaming_beam_equals_one_shot(seed, T, beam, chunk):
    """masr_ctc_prefix_beam_stream fed chunk by chunk (beam, trie and hash persist on the device) == one
    masr_ctc_prefix_beam call over all frames: bit-identical tokens and score after every chunk boundary's prefix."""
    import ctypes as C
    from masr_b200 import _lib
    _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    V = 4233
    _, logits = rand_posteriors(seed, T, V)
    ldl = (V + 15) // 16 * 16
    L = torch.zeros(T, ldl, device=dev)
    L[:, :V] = torch.from_numpy(logits).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    cid = torch.empty(T, 40, dtype=torch.int32, device=dev); clp = torch.empty(T, 40, device=dev)
    cn = torch.empty(T, dtype=torch.int32, device=dev)
    _lib.call("masr_ctc_topk_f32", L.data_ptr(), ldl, T, V, 40, 0.99, cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), st)
    pool_n, trie_n, si, sf = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _lib.call("masr_ctc_prefix_beam_workspace", 1, T, C.byref(pool_n), C.byref(trie_n))
    _lib.call("masr_ctc_prefix_beam_state_size", C.byref(si), C.byref(sf))

    def bufs():
        return (torch.empty(pool_n.value, device=dev), torch.empty(trie_n.value, dtype=torch.int32, device=dev),
                torch.empty(trie_n.value, dtype=torch.int32, device=dev), torch.zeros(1, T, dtype=torch.int32, device=dev),
                torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, device=dev))

    def one_shot(n):
        pool, tp, tt, otok, on, osc = bufs()
        ld = torch.tensor([n], dtype=torch.int32, device=dev)
        _lib.call("masr_ctc_prefix_beam", cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), T, ld.data_ptr(), 1, beam, 0, pool.data_ptr(),
                  tp.data_ptr(), tt.data_ptr(), trie_n.value, otok.data_ptr(), T, on.data_ptr(), osc.data_ptr(), st)
        return otok[0, :on.item()].cpu().tolist(), osc.item()

    pool, tp, tt, otok, on, osc = bufs()
    sti = torch.zeros(si.value, dtype=torch.int32, device=dev); stf = torch.zeros(sf.value, device=dev)
    done = 0
    while done < T:
        n = min(chunk, T - done)
        ld = torch.tensor([n], dtype=torch.int32, device=dev)
        _lib.call("masr_ctc_prefix_beam_stream", cid[done:].data_ptr(), clp[done:].data_ptr(), cn[done:].data_ptr(), T, ld.data_ptr(), 1,
                  beam, 0, pool.data_ptr(), tp.data_ptr(), tt.data_ptr(), trie_n.value, sti.data_ptr(), stf.data_ptr(),
                  1 if done else 0, otok.data_ptr(), T, on.data_ptr(), osc.data_ptr(), st)
        done += n
        want_t, want_s = one_shot(done)
        assert otok[0, :on.item()].cpu().tolist() == want_t, done
        assert osc.item() == want_s, done
    # a chunk without frames returns the standing result; resume = 0 starts over
    ld = torch.tensor([0], dtype=torch.int32, device=dev)
    _lib.call("masr_ctc_prefix_beam_stream", cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), T, ld.data_ptr(), 1, beam, 0, pool.data_ptr(),
              tp.data_ptr(), tt.data_ptr(), trie_n.value, sti.data_ptr(), stf.data_ptr(), 1, otok.data_ptr(), T, on.data_ptr(), osc.data_ptr(), st)
    assert (otok[0, :on.item()].cpu().tolist(), osc.item()) == one_shot(T)
    ld = torch.tensor([min(chunk, T)], dtype=torch.int32, device=dev)
    _lib.call("masr_ctc_prefix_beam_stream", cid.data_ptr(), clp.data_ptr(), cn.data_ptr(), T, ld.data_ptr(), 1, beam, 0, pool.data_ptr(),
              tp.data_ptr(), tt.data_ptr(), trie_n.value, sti.data_ptr(), stf.data_ptr(), 0, otok.data_ptr(), T, on.data_ptr(), osc.data_ptr(), st)
    assert (otok[0, :on.item()].cpu().tolist(), osc.item()) == one_shot(min(chunk, T))


@pytest.mark.gpu
def test_gpu_predict_stream_honours_ctc_beam_search(tmp_path):
    """`decoder: ctc_beam_search` + predict_stream (predict.py:320-322,352-353): after every push the result is the prefix
    beam search over all chunk posteriors so far — checked against the CPU restatement run on the engine's own chunk
    posteriors (one-shot over their concatenation), push by push; reset_stream starts a fresh search."""
    from conftest import make_audio, synth_weights
    from masr_b200 import synth
    from masr_b200.predict import MASRPredictor, chunk_starts, DECODING_WINDOW, CACHED_FEATURE_NUM
    mp, vp = str(tmp_path / "m.pt"), str(tmp_path / "vocabulary.txt")
    torch.save(synth.to_torch(synth_weights(0)), mp)
    synth.write_vocabulary(vp)
    base = {"use_model": "conformer", "streaming": True,
            "preprocess_conf": {"feature_method": "fbank", "n_mels": 80, "sample_rate": 16000, "use_dB_normalization": True, "target_dB": -20},
            "dataset_conf": {"dataset_vocab": vp}}
    beam_conf = {"beam_size": 300, "cutoff_prob": 0.99, "cutoff_top_n": 40}
    pred = MASRPredictor(configs={**base, "decoder": "ctc_beam_search", "ctc_beam_search_decoder_conf": beam_conf}, model_path=mp, use_gpu=True)
    x = make_audio("speech", 77, 16000 * 3 + 2000)
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    push = 8000
    got = [pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)]
    # the chunk posteriors of the same stream, from a second (greedy) predictor's engine, chunk by chunk
    eng = pred.predictor
    st = eng.new_stream()
    vocab = synth.vocabulary()
    remained, cached, probs_all, want = None, None, [], []
    from masr_b200.audio import pcm_bytes_to_float32
    for s in range(0, len(pcm), push):
        is_end = s + push >= len(pcm)
        new = pcm_bytes_to_float32(pcm[s:s + push].tobytes())
        remained = new if remained is None else np.concatenate([remained, new])
        feats, frames, _ = eng.fbank([remained])
        gain = float(eng.last_gain.cpu().numpy()[0])
        f = feats[0, :frames[0]]
        cached = f if cached is None else torch.cat([cached, f], 0)
        remained = (remained[160 * frames[0]:] * np.float32(gain)).astype(np.float32)
        starts = chunk_starts(int(cached.shape[0]), is_end)
        if not starts:
            want.append(None)
            continue
        for cur in starts:
            end = min(cur + DECODING_WINDOW, int(cached.shape[0]))
            out = eng.encode_chunk(cached[cur:end], st, required_cache_size=-16, want_probs=True)
            if out is not None:
                probs_all.append(out[2].cpu().numpy())
        cached = cached[end - CACHED_FEATURE_NUM:]
        (score, toks), = obeam.prefix_beam_search(np.concatenate(probs_all), **beam_conf)
        want.append({"text": "".join(vocab[i] for i in toks).replace("<space>", " "), "score": score})
    assert len(got) == len(want) and any(w is not None and w["text"] for w in want)
    for r, w in zip(got, want):
        assert (r is None) == (w is None)
        if r is not None:
            assert r["text"] == w["text"], (r, w)
            assert abs(r["score"] - w["score"]) < 5e-3 * max(1.0, abs(w["score"]))
    pred.reset_stream()
    again = [pred.predict_stream(audio_data=pcm[s:s + push].tobytes(), is_end=s + push >= len(pcm)) for s in range(0, len(pcm), push)]
    assert again == got
