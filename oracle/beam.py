"""Oracle (test infrastructure): CTC prefix beam search without a language model — PARITY UNPINNED.

The reference delegates ``ctc_beam_search`` to the external, un-vendored, absent ``paddlespeech_ctcdecoders`` SWIG/C++
library (masr/decoders/swig_wrapper.py:1,35-64; call sites masr/decoders/beam_search_decoder.py:45-56,75-91; version
unpinned: ``pip install paddlespeech_ctcdecoders -U``, docs/beam_search.md:5) plus a 2.8 GB KenLM file fetched at run
time (:19-25).  Neither the library, its source nor the LM exist in /root/reference or in this image, and no reference
test pins its results, so this restatement follows the algorithm's *public definition* (DeepSpeech2-style prefix beam
search) with the parameters MASR passes (beam_size=300, cutoff_prob=0.99, cutoff_top_n=40,
blank_id=0, configs/conformer.yml:74-88) and scorer=None.  It is validated only against itself: beam=1 on one-hot
posteriors == greedy, score monotonicity, and exact agreement with the CUDA implementation.

Definition used (log domain, natural log):
  per frame  : candidates = the tokens, sorted by probability (descending, ties by lower id), of the shortest prefix of
               that order whose cumulative mass >= cutoff_prob, capped at cutoff_top_n;  logp_c = log(p_c)
  per prefix : p_b (ends in blank), p_nb (ends in non-blank); root: p_b = 0, p_nb = -inf; score = p_b (+) p_nb
               blank c         : p_b'(l)   (+)= score(l) + logp_c
               c == last(l)    : p_nb'(l)  (+)= p_nb(l) + logp_c ;  p_nb'(l+c) (+)= p_b(l) + logp_c
               otherwise       : p_nb'(l+c) (+)= score(l) + logp_c
  prune      : keep the beam_size best prefixes by score' (ties: earlier-created prefix first)
  result     : prefixes best-first as (score, token ids); MASR takes element 0 (beam_search_decoder.py:56).
The library's early-exit heuristic (`min_cutoff`) is an optimisation and is not modelled.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

NEG_INF = -float("inf")


def logaddexp32(a: np.float32, b: np.float32) -> np.float32:
    """float32 log-sum-exp exactly as the CUDA kernel evaluates it: max + log1p(exp(min - max))."""
    a, b = np.float32(a), np.float32(b)
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    hi, lo = (a, b) if a >= b else (b, a)
    return np.float32(hi + np.float32(math.log1p(float(np.float32(math.exp(float(np.float32(lo - hi))))))))


def prune_frame(p: np.ndarray, cutoff_prob: float, cutoff_top_n: int) -> List[Tuple[int, np.float32]]:
    """-> [(token id, float32 probability)] kept for this frame, best first."""
    order = np.lexsort((np.arange(p.shape[0]), -p.astype(np.float64)))       # prob descending, id ascending on ties
    keep, cum = [], np.float32(0.0)
    for idx in order[:cutoff_top_n]:
        keep.append((int(idx), np.float32(p[idx])))
        cum = np.float32(cum + np.float32(p[idx]))
        if cum >= np.float32(cutoff_prob):
            break
    return keep


def prefix_beam_search(probs: np.ndarray, beam_size: int = 300, cutoff_prob: float = 0.99, cutoff_top_n: int = 40,
                       blank: int = 0, nbest: int = 1):
    """probs [T, V] float32 posteriors -> list of (score float, token id list), best first."""
    # a prefix is identified by a node id in a trie: node -> (parent node, last token); root = 0
    parent, last = [-1], [-1]
    child: Dict[Tuple[int, int], int] = {}
    beam = [(0, np.float32(0.0), np.float32(NEG_INF))]                       # (node, p_b, p_nb), best first
    for t in range(probs.shape[0]):
        cands = [(c, np.float32(math.log(float(pc)))) for c, pc in prune_frame(probs[t], cutoff_prob, cutoff_top_n)
                 if pc > 0]
        new_b: Dict[int, np.float32] = {}
        new_nb: Dict[int, np.float32] = {}
        order: List[int] = []                                                # creation / first-touch order for tie-breaks

        def touch(node):
            if node not in new_b:
                new_b[node], new_nb[node] = np.float32(NEG_INF), np.float32(NEG_INF)
                order.append(node)

        for node, pb, pnb in beam:                                           # existing prefixes keep their rank order
            touch(node)
        for node, pb, pnb in beam:
            score = logaddexp32(pb, pnb)
            for c, lp in cands:
                if c == blank:
                    new_b[node] = logaddexp32(new_b[node], np.float32(score + lp))
                    continue
                if c == last[node]:
                    new_nb[node] = logaddexp32(new_nb[node], np.float32(pnb + lp))
                    add = np.float32(pb + lp) if pb != NEG_INF else np.float32(NEG_INF)
                else:
                    add = np.float32(score + lp)
                if add == NEG_INF:
                    continue
                key = (node, c)
                ch = child.get(key)
                if ch is None:
                    ch = len(parent)
                    parent.append(node)
                    last.append(c)
                    child[key] = ch
                touch(ch)
                new_nb[ch] = logaddexp32(new_nb[ch], add)
        scored = []
        for rank, node in enumerate(order):
            s = logaddexp32(new_b[node], new_nb[node])
            if s != NEG_INF:
                scored.append((-float(s), rank, node))
        scored.sort()
        beam = [(node, new_b[node], new_nb[node]) for _, _, node in scored[:beam_size]]
    out = []
    for node, pb, pnb in beam[:nbest]:
        toks = []
        n = node
        while n > 0:
            toks.append(last[n])
            n = parent[n]
        out.append((float(logaddexp32(pb, pnb)), toks[::-1]))
    return out
