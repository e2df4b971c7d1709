"""Long-form recognition support: the segmentation half of ``MASRPredictor.predict_long``.

The reference runs the silero VAD network (an ONNX model shipped next to masr/infer_utils/vad_predictor.py, evaluated with
onnxruntime on the CPU, 512-sample windows) and turns its per-window speech probabilities into speech segments with a
hysteresis state machine (vad_predictor.py:106-175).  The network is a third-party model and stays what it is in the
reference — an ONNX session on the host (``SileroVAD``, needs ``onnxruntime`` and the model file; neither is part of this
image).  The state machine is restated here (``speech_timestamps_from_probs``) and pinned to the reference's own
implementation by tests/golden/vad_timestamps_golden.json; any object with the reference's
``get_speech_timestamps(samples, sampling_rate)`` method can be plugged into ``MASRPredictor.predict_long``.
The recognition half is where the GPU path changes the picture: all segments of a recording go through ONE batched pass
(``predict_batch``) instead of the reference's one-``predict``-per-segment loop (predict.py:216-224).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def speech_timestamps_from_probs(speech_probs: Sequence[float], audio_length_samples: int, sampling_rate: int = 16000,
                                 threshold: float = 0.5, min_speech_duration_ms: int = 250, min_silence_duration_ms: int = 100,
                                 window_size_samples: int = 512, speech_pad_ms: int = 30) -> List[Dict[str, int]]:
    """vad_predictor.py:114-175: per-window speech probabilities -> [{'start', 'end'}] in samples.

    A segment opens at the first window with p >= threshold, closes once p has stayed below threshold - 0.15 for
    ``min_silence_duration_ms`` (the close point is where the silence began), is kept if longer than
    ``min_speech_duration_ms``; afterwards segments are padded by ``speech_pad_ms`` (or share the gap when it is shorter
    than two pads)."""
    min_speech_samples = sampling_rate * min_speech_duration_ms / 1000
    min_silence_samples = sampling_rate * min_silence_duration_ms / 1000
    speech_pad_samples = sampling_rate * speech_pad_ms / 1000
    W = window_size_samples
    triggered = False
    speeches: List[Dict[str, int]] = []
    cur: Dict[str, int] = {}
    neg_threshold = threshold - 0.15
    temp_end = 0
    for i, p in enumerate(speech_probs):
        if p >= threshold and temp_end:
            temp_end = 0
        if p >= threshold and not triggered:
            triggered = True
            cur["start"] = W * i
            continue
        if p < neg_threshold and triggered:
            if not temp_end:
                temp_end = W * i
            if W * i - temp_end < min_silence_samples:
                continue
            cur["end"] = temp_end
            if cur["end"] - cur["start"] > min_speech_samples:
                speeches.append(cur)
            temp_end = 0
            cur = {}
            triggered = False
    if cur and (audio_length_samples - cur["start"]) > min_speech_samples:
        cur["end"] = audio_length_samples
        speeches.append(cur)
    for i, sp in enumerate(speeches):
        if i == 0:
            sp["start"] = int(max(0, sp["start"] - speech_pad_samples))
        if i != len(speeches) - 1:
            silence = speeches[i + 1]["start"] - sp["end"]
            if silence < 2 * speech_pad_samples:
                sp["end"] += int(silence // 2)
                speeches[i + 1]["start"] = int(max(0, speeches[i + 1]["start"] - silence // 2))
            else:
                sp["end"] = int(min(audio_length_samples, sp["end"] + speech_pad_samples))
                speeches[i + 1]["start"] = int(max(0, speeches[i + 1]["start"] - speech_pad_samples))
        else:
            sp["end"] = int(min(audio_length_samples, sp["end"] + speech_pad_samples))
    return speeches


class ProbabilityVAD:
    """Adapter: a callable ``window_probs(samples float32[n], sampling_rate) -> sequence of per-window speech probabilities``
    (one per ``window_size_samples`` window, the last one zero-padded) behind the reference's ``get_speech_timestamps``."""

    def __init__(self, window_probs, threshold: float = 0.5, min_speech_duration_ms: int = 250, min_silence_duration_ms: int = 100,
                 window_size_samples: int = 512, speech_pad_ms: int = 30):
        self.window_probs = window_probs
        self.kw = dict(threshold=threshold, min_speech_duration_ms=min_speech_duration_ms,
                       min_silence_duration_ms=min_silence_duration_ms, window_size_samples=window_size_samples,
                       speech_pad_ms=speech_pad_ms)

    def get_speech_timestamps(self, audio: np.ndarray, sampling_rate: int):
        probs = self.window_probs(np.asarray(audio, np.float32), sampling_rate)
        return speech_timestamps_from_probs(list(probs), len(audio), sampling_rate, **self.kw)


class SileroVAD(ProbabilityVAD):
    """The reference's VADPredictor (vad_predictor.py:11-104): the silero ONNX network on the host through onnxruntime,
    512-sample windows with the LSTM state carried across windows.  Needs ``onnxruntime`` and the model file."""

    def __init__(self, path: str, **kw):
        try:
            import onnxruntime
        except ImportError as e:                                    # not part of this image: fail loudly, no fallback
            raise RuntimeError("SileroVAD needs the `onnxruntime` package (the reference's VAD runs the silero ONNX model on "
                               "the CPU); pass another `vad_predictor` to predict_long or install it") from e
        self.session = onnxruntime.InferenceSession(path)
        super().__init__(self._probs, **kw)

    def _probs(self, audio: np.ndarray, sr: int):
        if sr != 16000 and sr % 16000 == 0:
            audio, sr = audio[::sr // 16000], 16000
        if sr not in (8000, 16000):
            raise ValueError("Supported sampling rates: [8000, 16000] (or multiply of 16000)")
        W = self.kw["window_size_samples"]
        h = np.zeros((2, 1, 64), np.float32)
        c = np.zeros((2, 1, 64), np.float32)
        out = []
        for s in range(0, len(audio), W):
            chunk = audio[s:s + W]
            if len(chunk) < W:
                chunk = np.pad(chunk, (0, W - len(chunk)))
            o, h, c = self.session.run(None, {"input": chunk[None].astype(np.float32), "h": h, "c": c,
                                              "sr": np.array(sr, dtype=np.int64)})
            out.append(float(np.asarray(o).item()))
        return out
