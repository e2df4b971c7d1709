"""Drop-in for ``masr.predict.MASRPredictor`` (masr/predict.py:19-362) on the B200 engine.

Same constructor arguments, same ``predict`` / ``predict_stream`` / ``reset_stream`` signatures,
same ``{'text': str, 'score': float}`` results, same YAML keys (``use_model``, ``streaming``,
``decoder``, ``preprocess_conf``, ``dataset_conf.dataset_vocab``) and the same ``inference.pt``
weights.  What changes is where the work happens: the waveform goes to the GPU once and only token
ids + a score come back (the reference featurises on the CPU, copies features up, copies the whole
[T,V] posterior down and decodes with numpy — predict.py:181-190, inference_predictor.py:59-64).

Additive entry points (the reference API is single-utterance): ``predict_batch``.
Out of the hot-path scope and therefore explicit errors here: punctuation (``use_pun``), inverse
text normalisation (``is_itn``), VAD long-audio segmentation (``predict_long``), model download
(``configs=None``), resampling, and ``use_gpu=False`` (there is no CPU path).
"""
from __future__ import annotations

import logging
import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import yaml

from . import SUPPORT_MODEL
from .audio import load_audio, pcm_bytes_to_float32, samples_to_float32
from .engine import ConformerEngine, EfficientConformerEngine, greedy_score, subsampled_len
from .text import TextFeaturizer, ids_to_text

logger = logging.getLogger(__name__)


class _Cfg(dict):
    """``dict_to_object`` (masr/utils/utils.py:45-56): attribute access over nested dicts."""
    __setattr__ = dict.__setitem__
    __getattr__ = dict.__getitem__


def dict_to_object(obj):
    if not isinstance(obj, dict):
        return obj
    out = _Cfg()
    for k, v in obj.items():
        out[k] = dict_to_object(v)
    return out


# streaming window arithmetic of predict.py:283-289
DECODING_CHUNK_SIZE = 16
CONTEXT = 7
SUBSAMPLING = 4
CACHED_FEATURE_NUM = CONTEXT - SUBSAMPLING                         # 3 feature frames carried over
DECODING_WINDOW = (DECODING_CHUNK_SIZE - 1) * SUBSAMPLING + CONTEXT  # 67
STRIDE = SUBSAMPLING * DECODING_CHUNK_SIZE                         # 64


def chunk_starts(num_frames: int, is_end: bool) -> List[int]:
    """Start indices of the encoder chunks the reference runs for ``num_frames`` cached feature
    frames (predict.py:292-303); empty when it would return ``None``."""
    if num_frames < DECODING_WINDOW and not is_end:
        return []
    if num_frames < CONTEXT:
        return []
    left = CONTEXT if is_end else DECODING_WINDOW
    return list(range(0, num_frames - left + 1, STRIDE))


class MASRPredictor:
    def __init__(self,
                 configs=None,
                 model_tag='conformer_streaming_fbank_aishell',
                 model_path='models/conformer_streaming_fbank/inference.pt',
                 use_pun=False,
                 pun_model_dir='models/pun_models/',
                 use_gpu=True):
        if not configs:
            raise Exception("masr_b200: model download (configs=None, model_tag=...) is not supported; "
                            "pass a YAML path or dict plus model_path")
        if isinstance(configs, str):
            with open(configs, 'r', encoding='utf-8') as f:
                configs = yaml.load(f.read(), Loader=yaml.FullLoader)
        self.configs = dict_to_object(configs)
        assert self.configs.use_model in SUPPORT_MODEL, f'没有该模型：{self.configs.use_model}'
        if not use_gpu:
            raise Exception("masr_b200 has no CPU path: use_gpu=False is not supported")
        if use_pun:
            raise Exception("masr_b200: the punctuation model (use_pun) is outside the hot-path scope")
        self.running = False
        self.use_gpu = use_gpu
        self._text_featurizer = TextFeaturizer(vocab_filepath=self.configs.dataset_conf.dataset_vocab)
        pc = self.configs.preprocess_conf
        if pc.get('feature_method', 'fbank') != 'fbank' or int(pc.get('n_mels', 80)) != 80:
            raise Exception("masr_b200 implements the fbank/80-mel front-end of the shipped configs only")
        self._sample_rate = int(pc.get('sample_rate', 16000))
        self._use_db = bool(pc.get('use_dB_normalization', True))
        self._target_db = float(pc.get('target_dB', -20))
        self._beam_conf = None
        if self.configs.decoder == 'ctc_beam_search':
            # GPU prefix beam search without a language model (the reference's external decoder + KenLM file are not
            # available; alpha/beta/language_model_path are ignored): whole-utterance calls (engine.ctc_beam) and streaming
            # (engine.StreamBeam = BeamSearchDecoder.decode_chunk / reset_decoder); both report the beam's log score.
            bc = dict(self.configs.get('ctc_beam_search_decoder_conf', {}) or {})
            self._beam_conf = {'beam_size': int(bc.get('beam_size', 300)), 'cutoff_prob': float(bc.get('cutoff_prob', 0.99)),
                               'cutoff_top_n': int(bc.get('cutoff_top_n', 40))}
            logger.warning('ctc_beam_search: GPU prefix beam search without LM (alpha/beta ignored)')
        if not os.path.exists(model_path):
            raise Exception("模型文件不存在，请检查{}是否存在！".format(model_path))
        from .squeezeformer import SqueezeformerEngine
        from .deepspeech2 import DeepSpeech2Engine
        engines = {'conformer': ConformerEngine, 'efficient_conformer': EfficientConformerEngine,
                   'squeezeformer': SqueezeformerEngine, 'deepspeech2': DeepSpeech2Engine}
        if self.configs.use_model not in engines:
            raise Exception(f"masr_b200: model '{self.configs.use_model}' is not implemented yet "
                            f"(available: {sorted(engines)})")
        self.predictor = engines[self.configs.use_model](model_path, streaming=bool(self.configs.streaming))
        self._can_stream = self.configs.use_model in ('conformer', 'deepspeech2', 'squeezeformer', 'efficient_conformer')
        if self.predictor.V != self._text_featurizer.vocab_size:
            raise Exception(f"vocabulary has {self._text_featurizer.vocab_size} entries but the model's CTC head has "
                            f"{self.predictor.V}")
        # streaming state (predict.py:70-73)
        self.remained_wav: Optional[np.ndarray] = None
        self.cached_feat: Optional[torch.Tensor] = None       # device [n, 80]
        self._stream = self.predictor.new_stream() if (self.configs.streaming and self._can_stream) else None
        self._hist_ids: List[int] = []
        self._hist_probs: List[np.float32] = []
        self._sbeam = None                                   # streaming beam-search state (created on the first chunk)
        self._sbeam_result = ([], 0.0)
        # warm-up, as the reference does (predict.py:88-93)
        warmup_audio = np.random.uniform(low=-2.0, high=2.0, size=(134240,))
        self.predict(audio_data=warmup_audio, is_itn=False)
        self.reset_stream()

    # ---------------------------------------------------------------------------------------------
    def _check_rate(self, sr):
        if sr != self._sample_rate:
            raise Exception(f"masr_b200: resampling is outside the hot-path scope (got {sr} Hz, model expects "
                            f"{self._sample_rate} Hz)")

    def _finish(self, text, use_pun, is_itn):
        if use_pun:
            logger.warning('标点符号模型没有初始化！')
        if is_itn:
            raise Exception("masr_b200: inverse text normalisation (is_itn) is outside the hot-path scope")
        return text

    def predict(self, audio_data, use_pun=False, is_itn=False, sample_rate=16000):
        """Whole-utterance recognition (predict.py:167-192)."""
        samples, sr = load_audio(audio_data, sample_rate)
        self._check_rate(sr)
        if self._beam_conf is not None:
            toks, scores = self.predictor.transcribe_beam([samples], use_db_normalization=self._use_db,
                                                          target_db=self._target_db, **self._beam_conf)
            text = ids_to_text(toks[0], self._text_featurizer.vocab_list)
            return {'text': self._finish(text, use_pun, is_itn), 'score': scores[0]}
        res = self.predictor.transcribe([samples], self._use_db, self._target_db)
        self._raise_status(res.status)
        text = ids_to_text(res.tokens[0], self._text_featurizer.vocab_list)
        return {'text': self._finish(text, use_pun, is_itn), 'score': res.scores[0]}

    def predict_batch(self, audio_list: Sequence, sample_rate=16000):
        """Additive: a list of utterances in one GPU pass; element i equals ``predict(audio_list[i])``."""
        waves = []
        for a in audio_list:
            s, sr = load_audio(a, sample_rate)
            self._check_rate(sr)
            waves.append(s)
        if self._beam_conf is not None:
            toks, scores = self.predictor.transcribe_beam(waves, use_db_normalization=self._use_db,
                                                          target_db=self._target_db, **self._beam_conf)
            vocab = self._text_featurizer.vocab_list
            return [{'text': ids_to_text(t, vocab), 'score': s} for t, s in zip(toks, scores)]
        res = self.predictor.transcribe(waves, self._use_db, self._target_db)
        self._raise_status(res.status)
        vocab = self._text_featurizer.vocab_list
        return [{'text': ids_to_text(t, vocab), 'score': s} for t, s in zip(res.tokens, res.scores)]

    def predict_batches(self, batches, sample_rate=16000, device_hook=None):
        """Additive: a stream of batches (iterable of lists of utterances) -> one list of ``{'text','score'}`` per batch, in
        order; element i of batch k equals ``predict(batches[k][i])``.  Host staging and the H2D copy of batch k+1 overlap
        the GPU pass of batch k (``ConformerEngine.transcribe_pipelined``), so the results lag the input by one batch.
        Greedy decoding only.  ``device_hook``: see ``transcribe_pipelined`` (cross-rank gather of a sharded deployment)."""
        vocab = self._text_featurizer.vocab_list
        if self._beam_conf is not None:
            # the prefix beam search of batch k runs on a second stream under the encoder of batch k+1
            def loaded_b():
                for audio_list in batches:
                    waves = []
                    for a in audio_list:
                        s, sr = load_audio(a, sample_rate)
                        self._check_rate(sr)
                        waves.append(s)
                    yield waves
            for toks, scores in self.predictor.transcribe_beam_pipelined(loaded_b(), use_db_normalization=self._use_db,
                                                                         target_db=self._target_db, **self._beam_conf):
                yield [{'text': ids_to_text(t, vocab), 'score': s} for t, s in zip(toks, scores)]
            return

        def loaded():
            for audio_list in batches:
                waves = []
                for a in audio_list:
                    s, sr = load_audio(a, sample_rate)
                    self._check_rate(sr)
                    waves.append(s)
                yield waves
        for res in self.predictor.transcribe_pipelined(loaded(), self._use_db, self._target_db, device_hook=device_hook):
            self._raise_status(res.status)
            yield [{'text': ids_to_text(t, vocab), 'score': s} for t, s in zip(res.tokens, res.scores)]

    @staticmethod
    def _raise_status(status):
        if status is not None and np.any(status != 0):
            # AudioSegment.normalize raises ValueError when gain > max_gain_db (audio.py:301-303)
            raise ValueError("无法将段规范化到目标dB，音频增益已经超过max_gain_db (300.0dB)")

    def init_vad(self, vad_predictor=None, vad_model_path=None):
        """predict.py:139-142.  ``vad_predictor``: any object with the reference's ``get_speech_timestamps(samples,
        sampling_rate)``; else the silero ONNX model at ``vad_model_path`` through onnxruntime (masr_b200.vad.SileroVAD)."""
        if vad_predictor is not None:
            self.vad_predictor = vad_predictor
        elif getattr(self, "vad_predictor", None) is None:
            if vad_model_path is None:
                raise Exception("masr_b200: predict_long needs a VAD: pass vad_predictor=... (an object with "
                                "get_speech_timestamps) or vad_model_path=<silero_vad.onnx> (needs onnxruntime)")
            from .vad import SileroVAD
            self.vad_predictor = SileroVAD(vad_model_path)

    def predict_long(self, audio_data, use_pun=False, is_itn=False, sample_rate=16000, vad_predictor=None, vad_model_path=None):
        """Long-form recognition (predict.py:195-234): VAD segments -> recognise -> join with '，' and average the scores.
        All segments of the recording go through ONE batched GPU pass (``predict_batch``) instead of the reference's
        one-``predict``-per-segment loop; each segment's result equals ``predict(segment)`` (B=1 semantics)."""
        self.init_vad(vad_predictor, vad_model_path)
        samples, sr = load_audio(audio_data, sample_rate)
        self._check_rate(sr)
        stamps = self.vad_predictor.get_speech_timestamps(samples, sr)
        segs = [samples[t['start']:t['end']] for t in stamps]
        results = self.predict_batch(segs, sample_rate=sr) if segs else []
        texts, scores = '', []
        for r in results:
            if r['text'] != '':
                texts = texts + r['text'] if use_pun else texts + '，' + r['text']
            scores.append(r['score'])
        if texts[:1] == '，':
            texts = texts[1:]
        if use_pun and len(texts) > 0:
            logger.warning('标点符号模型没有初始化！')
        if is_itn:
            raise Exception("masr_b200: inverse text normalisation (is_itn) is outside the hot-path scope")
        return {'text': texts, 'score': round(sum(scores) / len(scores), 2) if scores else 0}

    # ---------------------------------------------------------------------------------------------
    def predict_stream(self, audio_data, is_end=False, use_pun=False, is_itn=False, channels=1, samp_width=2,
                       sample_rate=16000):
        """Streaming recognition, one push of audio per call (predict.py:237-343).  Returns ``None``
        while fewer than 67 feature frames are buffered, else the running ``{'text','score'}``."""
        if not self.configs.streaming:
            raise Exception(
                f"不支持改该模型流式识别，当前模型：{self.configs.use_model}，参数streaming为：{self.configs.streaming}")
        if not self._can_stream:
            raise NotImplementedError(f"masr_b200: predict_stream is not implemented for '{self.configs.use_model}' yet")
        if isinstance(audio_data, np.ndarray):
            new = samples_to_float32(audio_data)
        elif isinstance(audio_data, bytes):
            new = pcm_bytes_to_float32(audio_data, channels=channels, samp_width=samp_width)
        else:
            raise Exception(f'不支持该数据类型，当前数据类型为：{type(audio_data)}')
        self._check_rate(sample_rate)
        self.remained_wav = new if self.remained_wav is None else np.concatenate([self.remained_wav, new])

        # featurise everything not yet consumed; the reference dB-normalises the remainder IN PLACE on
        # every push (predict.py:274 + audio_featurizer.py:49-50), so the carried-over tail keeps the gain
        eng = self.predictor
        feats, frames, status = eng.fbank([self.remained_wav], self._use_db, self._target_db)
        gain = float(eng.last_gain.cpu().numpy()[0]) if self._use_db else 1.0
        self._raise_status(status.cpu().numpy())
        nf = frames[0]
        x_chunk = feats[0, :nf]
        self.cached_feat = x_chunk if self.cached_feat is None else torch.cat([self.cached_feat, x_chunk], dim=0)
        tail = self.remained_wav[FRAME_SHIFT * nf:]
        self.remained_wav = (tail * np.float32(gain)).astype(np.float32) if self._use_db else tail

        num_frames = int(self.cached_feat.shape[0])
        starts = chunk_starts(num_frames, is_end)
        if not starts:
            return None
        end = None
        for cur in starts:
            end = min(cur + DECODING_WINDOW, num_frames)
            out = eng.encode_chunk(self.cached_feat[cur:end], self._stream,
                                   required_cache_size=DECODING_CHUNK_SIZE * -1)
            if out is None:
                continue
            ids, maxp, _ = out
            if self._beam_conf is not None:
                # predict.py:320-322: beam_search_decoder.decode_chunk on this chunk's posteriors (state kept on the device)
                if self._sbeam is None:
                    from .engine import StreamBeam
                    self._sbeam = StreamBeam(eng, **self._beam_conf)
                self._sbeam_result = self._sbeam.push(self._stream.last_logits, int(ids.shape[0]))
                continue
            ids_h = ids.cpu().numpy()
            mp_h = maxp.cpu().numpy()
            self._hist_ids.extend(int(i) for i in ids_h)
            self._hist_probs.extend(mp_h[t] for t in range(len(ids_h)) if ids_h[t] != 0)
        self.cached_feat = self.cached_feat[end - CACHED_FEATURE_NUM:]
        if self._beam_conf is not None:
            toks, score = self._sbeam_result
            if is_itn:
                raise Exception("masr_b200: inverse text normalisation (is_itn) is outside the hot-path scope")
            return {'text': ids_to_text(toks, self._text_featurizer.vocab_list), 'score': score}
        # greedy_decoder_chunk re-collapses the whole history (ctc_greedy_decoder.py:81-88)
        toks, prev = [], None
        for i in self._hist_ids:
            if i != prev and i != 0:
                toks.append(i)
            prev = i
        acc = np.float32(0.0)
        for p in self._hist_probs:
            acc = np.float32(acc + p)
        score = greedy_score(acc, len(self._hist_probs))
        text = ids_to_text(toks, self._text_featurizer.vocab_list)
        if use_pun and is_end and len(text) > 0:
            logger.warning('标点符号模型没有初始化！')
        if is_itn:
            raise Exception("masr_b200: inverse text normalisation (is_itn) is outside the hot-path scope")
        return {'text': text, 'score': score}

    def reset_stream(self):
        """predict.py:346-353."""
        if self._stream is not None:
            self._stream.reset()
        self.remained_wav = None
        self.cached_feat = None
        self._hist_ids = []
        self._hist_probs = []
        if self._sbeam is not None:                          # predict.py:352-353: beam_search_decoder.reset_decoder()
            self._sbeam.reset()
        self._sbeam_result = ([], 0.0)


FRAME_SHIFT = 160
