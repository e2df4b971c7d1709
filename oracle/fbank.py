"""Oracle (test infrastructure): the reference's audio front-end, restated in numpy float32.

Follows, step by step:
  * ``AudioSegment._convert_samples_to_float32``      masr/data_utils/audio.py:532-546
  * ``AudioSegment.rms_db`` / ``normalize`` / ``gain_db`` masr/data_utils/audio.py:519-529,287-304,256-264
  * ``AudioSegment._convert_samples_from_float32``     masr/data_utils/audio.py:549-574
  * ``AudioFeaturizer.featurize`` / ``_compute_fbank``   masr/data_utils/featurizer/audio_featurizer.py:37-69,120-138
  * ``torchaudio.compliance.kaldi.fbank`` (un-vendored third-party arithmetic; the reference
    leaves torchaudio unpinned, docs/install.md:7 names 2.0.2; the build container has 2.11.0):
    ``_get_strided`` kaldi.py:44-83, ``_get_window`` :154-217, ``get_mel_banks`` :436-511,
    ``fbank`` :514-645.
Pinned against the real ``AudioFeaturizer`` + torchaudio in tests/test_oracle_pinned.py and
tests/golden/fbank_*.npz.
"""
import math

import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400      # 25 ms  (audio_featurizer.py:125; kaldi.py:141)
FRAME_SHIFT = 160    # 10 ms
NFFT = 512           # round_to_power_of_two (kaldi.py:142)
NUM_MEL = 80
PREEMPH = 0.97
LOW_FREQ = 20.0
EPS = np.float32(1.1920928955078125e-07)  # torch.finfo(float32).eps (kaldi.py:18)


def to_float32(samples: np.ndarray) -> np.ndarray:
    """audio.py:24-32,532-546 — ints scaled by 2^-(bits-1); multi-channel -> channel mean."""
    out = samples.astype(np.float32)
    if samples.dtype in (np.int8, np.int16, np.int32, np.int64):
        out *= np.float32(1.0 / 2 ** (np.iinfo(samples.dtype).bits - 1))
    elif samples.dtype not in (np.float16, np.float32, np.float64):
        raise TypeError("Unsupported sample type: %s." % samples.dtype)
    if out.ndim >= 2:
        out = np.mean(out, 1)
    return out


def pcm_bytes_to_float32(buf: bytes, samp_width: int = 2) -> np.ndarray:
    """masr/data_utils/utils.py:382-411 (``buf_to_float``)."""
    scale = 1.0 / float(1 << (8 * samp_width - 1))
    return scale * np.frombuffer(buf, "<i%d" % samp_width).astype(np.float32)


def normalize_gain(x: np.ndarray, target_db: float = -20.0, max_gain_db: float = 300.0):
    """Return ``(gained copy, gain factor)``; audio.py:287-304 with rms_db :519-529.
    The reference does this in place on the float32 sample buffer."""
    ms = np.mean(x ** 2)                       # float32 pairwise sum / n
    if ms == 0:
        ms = 1
    rms_db = 10 * np.log10(ms)
    gain = target_db - rms_db
    if gain > max_gain_db:
        raise ValueError("gain %r dB exceeds max_gain_db" % float(gain))
    factor = 10. ** (min(max_gain_db, gain) / 20.)
    y = x.copy()
    y *= factor
    return y, np.float32(factor)


def to_int16(x: np.ndarray) -> np.ndarray:
    """audio.py:549-574 — scale by 2^15, clip to the int16 range, truncate toward zero."""
    y = x.copy()
    y *= np.float32(32768.0)
    y[y > 32767] = 32767
    y[y < -32768] = -32768
    return y.astype(np.int16)


def povey_window() -> np.ndarray:
    """kaldi.py:100 — hann(400, periodic=False) ** 0.85."""
    n = np.arange(FRAME_LEN, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / (FRAME_LEN - 1))
    return (hann.astype(np.float32) ** np.float32(0.85)).astype(np.float32)


def mel_scale(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def mel_banks() -> np.ndarray:
    """kaldi.py:436-511 -> float32 [80, 257] (last column, Nyquist, is the zero pad of :627)."""
    nyq = 0.5 * SAMPLE_RATE
    fft_bin_width = SAMPLE_RATE / NFFT
    mlo, mhi = mel_scale(LOW_FREQ), mel_scale(nyq)
    delta = (mhi - mlo) / (NUM_MEL + 1)
    b = np.arange(NUM_MEL, dtype=np.float32)[:, None]
    left = np.float32(mlo) + b * np.float32(delta)
    center = np.float32(mlo) + (b + 1.0) * np.float32(delta)
    right = np.float32(mlo) + (b + 2.0) * np.float32(delta)
    mel = mel_scale(np.float32(fft_bin_width) * np.arange(NFFT // 2, dtype=np.float32))[None, :].astype(np.float32)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = np.maximum(np.float32(0), np.minimum(up, down)).astype(np.float32)
    return np.concatenate([w, np.zeros((NUM_MEL, 1), np.float32)], axis=1)


def num_frames(num_samples: int) -> int:
    """kaldi.py:63-67 (snip_edges=True)."""
    if num_samples < FRAME_LEN:
        return 0
    return 1 + (num_samples - FRAME_LEN) // FRAME_SHIFT


def kaldi_fbank(wave_i16: np.ndarray) -> np.ndarray:
    """int16 samples (as the reference passes them, ``.float()``-ed) -> float32 [F, 80] log-mel."""
    x = wave_i16.astype(np.float32)
    F = num_frames(x.shape[0])
    if F == 0:
        return np.zeros((0, NUM_MEL), np.float32)
    idx = np.arange(FRAME_LEN)[None, :] + FRAME_SHIFT * np.arange(F)[:, None]
    fr = x[idx]                                                    # kaldi.py:82
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=np.float32)     # :183-186
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)         # replicate-left pad :195-197
    fr = fr - np.float32(PREEMPH) * prev                           # :198
    fr = fr * povey_window()[None, :]                              # :201-204
    pad = np.zeros((F, NFFT - FRAME_LEN), np.float32)
    fr = np.concatenate([fr, pad], axis=1)                         # :207-211
    spec = np.fft.rfft(fr, axis=1)
    mag = np.abs(spec).astype(np.float32)                          # :616
    power = mag * mag                                              # :618
    e = power @ mel_banks().T                                      # :630
    return np.log(np.maximum(e, EPS)).astype(np.float32)           # :633


def featurize(samples: np.ndarray, use_db_normalization: bool = True, target_db: float = -20.0) -> np.ndarray:
    """``AudioFeaturizer.featurize`` on an array the way ``AudioSegment.from_ndarray`` sees it
    (audio_featurizer.py:37-69): float32 in [-1,1) -> dB normalise -> int16 -> Kaldi fbank."""
    x = to_float32(np.asarray(samples))
    if use_db_normalization:
        x, _ = normalize_gain(x, target_db)
    return kaldi_fbank(to_int16(x))
