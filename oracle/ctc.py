"""Oracle (test infrastructure): CTC greedy (best-path) decoding, restated in numpy.

Follows masr/decoders/ctc_greedy_decoder.py:
  * ``greedy_decoder``        :6-31   — first-max argmax over *probabilities*, score = 100 * mean of the
                                        max-probs of the non-blank frames (sequential float32 sum), collapse
                                        consecutive repeats, drop blank(0), ``<space>`` -> ' '.
  * ``greedy_decoder_chunk``  :52-89  — the streaming variant re-collapses the whole id history.  (The
                                        reference's two history lists have swapped names, :78-79; behaviour
                                        is what is restated here.)
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np


def best_path(probs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """probs [T,V] -> (ids int64 [T], max-prob float32 [T]); ``np.argmax`` = lowest index on ties."""
    probs = np.asarray(probs)
    ids = probs.argmax(axis=1)
    return ids, probs[np.arange(probs.shape[0]), ids]


def collapse(ids: Sequence[int], blank: int = 0) -> List[int]:
    out, prev = [], None
    for i in ids:
        i = int(i)
        if i != prev and i != blank:
            out.append(i)
        prev = i
    return out


def score_of(max_probs: Sequence[np.float32]) -> float:
    """``float(sum(list)/len(list)) * 100.0`` with numpy float32 scalars: a left-to-right float32 sum."""
    if len(max_probs) == 0:
        return 0
    acc = np.float32(0.0)
    for p in max_probs:
        acc = np.float32(acc + np.float32(p))
    return float(np.float32(acc / np.float32(len(max_probs)))) * 100.0


def ids_to_text(ids: Sequence[int], vocabulary: Sequence[str]) -> str:
    return "".join(vocabulary[i] for i in ids).replace("<space>", " ")


def greedy_decode(probs: np.ndarray, vocabulary: Sequence[str], blank: int = 0):
    """-> (score, text, collapsed ids)."""
    ids, mp = best_path(probs)
    kept = [mp[t] for t in range(len(ids)) if ids[t] != blank]
    tokens = collapse(ids, blank)
    return score_of(kept), ids_to_text(tokens, vocabulary), tokens


class GreedyStream:
    """State carried by ``MASRPredictor`` between ``greedy_decoder_chunk`` calls (predict.py:72-73,325-328)."""

    def __init__(self):
        self.ids: List[int] = []
        self.kept_probs: List[np.float32] = []

    def push(self, probs: np.ndarray, vocabulary: Sequence[str], blank: int = 0):
        ids, mp = best_path(probs)
        self.ids.extend(int(i) for i in ids)
        self.kept_probs.extend(mp[t] for t in range(len(ids)) if ids[t] != blank)
        tokens = collapse(self.ids, blank)
        return score_of(self.kept_probs), ids_to_text(tokens, vocabulary), tokens
