"""Many live streams on one GPU: batched chunk decoding for the streaming Conformer, Squeezeformer and EfficientConformer.

The reference's streaming API is one stream per ``MASRPredictor`` (predict.py:237-343; ``forward_chunk`` asserts
batch 1, conformer/encoder.py:378) and `infer_server.py` effectively serves one stream at a time.  BASELINE.json
config 3/5 ask for dozens to hundreds of concurrent streams.  ``ConformerStreamPool`` keeps N slots of streaming state
(attention K|V caches as fp16 pairs, conv left contexts, offsets) in one set of device buffers and advances every slot
that has a chunk ready in ONE pass of tensor-core GEMMs (M = slots x 16 rows) — each slot computed exactly like the
single-stream ``encode_chunk`` / reference ``forward_chunk`` with ``required_cache_size < 0`` (all history kept, which
is what ``predict_stream`` passes).  ``StreamPool`` adds the per-stream host logic of ``predict_stream`` (sample
carry-over with in-place dB renormalisation, 67/64/3 feature windowing, greedy history).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ._lib import EPI_BIAS, EPI_BIAS_GLU, EPI_BIAS_SCALE, EPI_BIAS_SILU, EPI_RESIDUAL
from .audio import pcm_bytes_to_float32, samples_to_float32
from .engine import ConformerEngine, _p, greedy_score, subsampled_len
from .predict import CACHED_FEATURE_NUM, DECODING_WINDOW, FRAME_SHIFT, chunk_starts
from .text import ids_to_text

CHUNK_FRAMES = DECODING_WINDOW          # 67 feature frames -> 16 encoder frames
CHUNK_OUT = 16


class _PoolBase:
    """Shared machinery of the batched chunk-decoding pools.

    Everything that varies from step to step is DEVICE DATA (one int32 table `meta`, refreshed by a single small H2D copy):
    per slot the valid query rows, key rows and cache fill at the full and at the halved frame rate.  The launch sequence
    of a step is therefore fixed, and after one eager step it is captured into a CUDA graph and replayed (a chunk step is
    ~150-250 small launches: launch-bound otherwise).  K|V rows are appended and conv left contexts slid by two
    bookkeeping kernels (csrc/stream.cu) instead of host-built index lists."""

    # rows of `meta`
    QLEN, KLEN, BASE, QLEN2, KLEN2, BASE2 = range(6)
    OUT_ROWS = CHUNK_OUT          # output frames per slot and chunk (8 for the EfficientConformer)

    def _init_common(self, eng, n_slots: int, use_graph: bool, keep_probs: bool = False):
        self.eng, self.S = eng, n_slots
        # keep_probs: also write the CTC posteriors of every output frame ([S * OUT_ROWS, V], `self.probs`) — the
        # `InferencePredictor.predict_chunk_conformer` seam (inference_predictor.py:80-94) returns them
        self.probs = (torch.zeros(n_slots * self.OUT_ROWS, eng.V, device=eng.device, dtype=torch.float32) if keep_probs else None)
        dev = eng.device
        self.meta = torch.zeros(6, n_slots, device=dev, dtype=torch.int32)
        self.meta_host = torch.zeros(6, n_slots, dtype=torch.int32, pin_memory=True)
        self.feats_in = torch.zeros(n_slots, CHUNK_FRAMES, 80, device=dev, dtype=torch.float32)
        self.lens_host = [0] * n_slots
        self.use_graph = bool(use_graph) and eng.use_graphs
        # LayerNorm folded into the preceding residual projection (cluster kernel, 4 launches per block fewer): measured r02
        # SLOWER here too (64 streams: 3.77 vs 3.51-3.59 ms per push, 547 vs 710 launches) — the cluster launch + DSMEM exchange
        # cost more than the ~2 us LayerNorm launch they replace at M = 1024.  Off unless MASR_POOL_FUSE_LN=1 (tests cover both).
        import os
        self.fuse_ln = os.environ.get("MASR_POOL_FUSE_LN", "0") == "1" and eng.d == 256
        self._graph = None
        self._graph_launches = 0
        self._warm = False
        self._meta_ev = None          # the previous step's H2D copy out of `meta_host` (pinned) has completed

    def _m(self, row):
        return self.meta[row].data_ptr()

    def _prepare(self, nframes: Sequence[int], short_ok_once: bool):
        eng, S, C = self.eng, self.S, CHUNK_OUT
        tout = [subsampled_len(int(n)) for n in nframes]
        tout2 = [(t + 1) // 2 for t in tout]
        for s in range(S):
            if short_ok_once and tout[s] and self.lens_host[s] % C:
                raise AssertionError(f"stream slot {s}: a short (final) chunk was already decoded; reset the stream first")
            if self.lens_host[s] + tout[s] > self.cap or self.lens_host[s] + tout[s] >= eng.w.max_len:
                raise AssertionError(f"stream slot {s}: {self.lens_host[s] + tout[s]} cached frames exceed the pool capacity")
        if self._meta_ev is not None:
            self._meta_ev.synchronize()
        mh = self.meta_host.numpy()
        mh[self.QLEN] = tout
        mh[self.BASE] = self.lens_host
        mh[self.KLEN] = mh[self.BASE] + mh[self.QLEN]
        mh[self.QLEN2] = tout2
        mh[self.BASE2] = mh[self.BASE] // 2
        mh[self.KLEN2] = mh[self.BASE2] + mh[self.QLEN2]
        self.meta.copy_(self.meta_host, non_blocking=True)
        self._meta_ev = torch.cuda.Event()
        self._meta_ev.record(torch.cuda.current_stream(eng.device))
        return tout, tout2

    def _run(self):
        """Eager on the first step (one-time setup: function attributes, tables), then capture once and replay."""
        eng = self.eng
        if not self.use_graph:
            self._body()
            return
        if not self._warm:
            self._body()
            self._warm = True
            return
        if self._graph is None:
            n0 = eng.launches
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._body()
            self._graph_launches = eng.launches - n0
            eng.launches = n0
            self._graph = g
        self._graph.replay()
        eng.launches += self._graph_launches

    def _append_pair(self, src, ld_src_elems, col0_elems, ncols, dst, cap, base_row, cnt_row, rows_per_slot, elem_bytes=2):
        """dst pair rows (s*cap + base[s] + t) <- src pair rows (s*rows_per_slot + t), columns [col0, col0+ncols)."""
        s1, d1 = (None, None) if not isinstance(src, tuple) else (src[1], dst[1])
        s0, d0 = (src, dst) if not isinstance(src, tuple) else (src[0], dst[0])
        self.eng._k("stream_append", "masr_stream_append_rows", _p(s0), _p(s1), ld_src_elems * elem_bytes, col0_elems * elem_bytes,
                    ncols * elem_bytes, _p(d0), _p(d1), d0.shape[-1] * elem_bytes, cap, self._m(base_row), self._m(cnt_row),
                    rows_per_slot, self.S)

    def _shift(self, x0, x1, rows_per_slot, lorder, row_bytes, cnt_row):
        self.eng._k("stream_shift", "masr_stream_shift_cache", _p(x0), _p(x1), rows_per_slot, lorder, row_bytes, self._m(cnt_row), self.S)

    def step(self, feats: torch.Tensor, nframes: Sequence[int]):
        """feats [S, 67, 80] raw log-mel (device), nframes[s] = valid feature frames of slot s this round (0 = idle).
        -> (ids [S, OUT_ROWS] int32, maxp [S, OUT_ROWS], valid output frames per slot)."""
        tout, tout2 = self._prepare(nframes, self.SHORT_ONCE)
        if feats.data_ptr() != self.feats_in.data_ptr():
            self.feats_in.copy_(feats)
        self._run()
        for s in range(self.S):
            self.lens_host[s] += tout[s]
        R = self.OUT_ROWS
        return self.b["ids"].view(self.S, R), self.b["maxp"].view(self.S, R), (tout if R == CHUNK_OUT else tout2)


class ConformerStreamPool(_PoolBase):
    """Device state + one batched chunk step for `n_slots` streams."""

    SHORT_ONCE = False            # every block runs at the full frame rate: short chunks may be followed by more chunks

    def __init__(self, eng: ConformerEngine, n_slots: int, max_frames: int = 3000, use_graph: bool = True, keep_probs: bool = False):
        if not eng.causal:
            raise Exception("chunk decoding needs a streaming (causal) model")
        if eng.gemm_path != "tc":
            raise Exception("the stream pool runs on the tensor-core path")
        self._init_common(eng, n_slots, use_graph, keep_probs)
        self.cap = max_frames
        dev, d, w = eng.device, eng.d, eng.w
        f16, f32 = torch.float16, torch.float32
        nl = len(w.layers)
        S, C = n_slots, CHUNK_OUT
        self.lorder = w.kernel - 1
        F1 = (CHUNK_FRAMES - 1) // 2
        TH = (F1 + 1) // 2
        M = S * C
        self.kv = [(torch.zeros(S * self.cap, 2 * d, device=dev, dtype=f16), torch.zeros(S * self.cap, 2 * d, device=dev, dtype=f16))
                   for _ in range(nl)]
        self.xcat = torch.zeros(nl, S, self.lorder + C, d, device=dev, dtype=f32)
        self.b = {
            "c1p": (torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16), torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16)),
            "c2p": (torch.empty(M * eng.f2, d, device=dev, dtype=f16), torch.empty(M * eng.f2, d, device=dev, dtype=f16)),
            "x": torch.empty(M, d, device=dev, dtype=f32), "t0": torch.empty(M, d, device=dev, dtype=f32),
            "t0p": (torch.empty(M, d, device=dev, dtype=f16), torch.empty(M, d, device=dev, dtype=f16)),
            "t1p": (torch.empty(M, d, device=dev, dtype=f16), torch.empty(M, d, device=dev, dtype=f16)),
            "hidp": (torch.empty(M, w.ffn, device=dev, dtype=f16), torch.empty(M, w.ffn, device=dev, dtype=f16)),
            "qkv": torch.empty(M, 3 * d, device=dev, dtype=f32),
            "qkvp": (torch.empty(M, 3 * d, device=dev, dtype=f16), torch.empty(M, 3 * d, device=dev, dtype=f16)),
            "xcp": (torch.empty(S * (self.lorder + C), d, device=dev, dtype=f16), torch.empty(S * (self.lorder + C), d, device=dev, dtype=f16)),
            "g": torch.empty(S * (self.lorder + C), d, device=dev, dtype=f32),
            "logits": torch.empty(M, eng.Vpad, device=dev, dtype=f32),
            "ids": torch.empty(M, device=dev, dtype=torch.int32), "maxp": torch.empty(M, device=dev, dtype=f32),
            "clen": torch.full((S,), self.lorder + C, device=dev, dtype=torch.int32),
        }

    def reset(self, slot: int):
        """``InferencePredictor.reset_stream`` for one slot (inference_predictor.py:97-102)."""
        self.lens_host[slot] = 0
        self.xcat[:, slot].zero_()

    def _body(self):
        """One batched ``forward_chunk`` over all slots (fixed launch sequence; per-slot lengths come from `meta`)."""
        eng, S, C, cap = self.eng, self.S, CHUNK_OUT, self.cap
        w, d, tw, b = eng.w, eng.d, eng._tcw, self.b
        feats = self.feats_in
        qlen, klen = self._m(self.QLEN), self._m(self.KLEN)
        M = S * C
        F1 = (CHUNK_FRAMES - 1) // 2
        x, t0, t0p, t1p, hidp, qkv, qkvp, g, xcp = b["x"], b["t0"], b["t0p"], b["t1p"], b["hidp"], b["qkv"], b["qkvp"], b["g"], b["xcp"]
        eng._k("conv1", "masr_conv1_cmvn_relu_planes_f16", _p(feats), _p(w.cmvn_mean), _p(w.cmvn_istd), _p(w.conv1_w), _p(w.conv1_b),
               _p(b["c1p"][0]), _p(b["c1p"][1]), S, CHUNK_FRAMES, w.idim, F1, eng.w1_cols, d)
        eng._k("conv2", "masr_conv2_tc_f16x2", _p(b["c1p"][0]), _p(b["c1p"][1]), _p(tw["conv2"][0]), _p(tw["conv2"][1]), _p(w.conv2_b),
               None, _p(b["c2p"][0]), _p(b["c2p"][1]), S, F1, C, d)
        eng._tc(b["c2p"], eng.f2 * d, tw["embed"], w.embed_b, M, d, eng.f2 * d, EPI_BIAS_SCALE, float(d) ** 0.5, C=x, ldc=d)
        LC = self.lorder + C
        fz, nl = self.fuse_ln, len(w.layers)
        for i, L in enumerate(w.layers):
            if not fz or i == 0:
                eng._ln_split(x, L.ln_ffm, t0p, M)
            eng._tc(t0p, d, tw[i, "ffm1"], L.ffm[1], M, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            if fz:
                eng._tc_ln(hidp, w.ffn, tw[i, "ffm2"], L.ffm[3], M, w.ffn, 0.5, x, L.ln_mha, t0p)
            else:
                eng._tc(hidp, w.ffn, tw[i, "ffm2"], L.ffm[3], M, d, w.ffn, EPI_RESIDUAL, 0.5, x, d, C=x, ldc=d)
                eng._ln_split(x, L.ln_mha, t0p, M)
            eng._tc(t0p, d, tw[i, "qkv"], L.bqkv, M, 3 * d, d, C=qkv, Cp=qkvp, ldc=3 * d)
            kvh, kvl = self.kv[i]
            self._append_pair(qkvp, 3 * d, d, 2 * d, (kvh, kvl), cap, self.BASE, self.QLEN, C)      # new K|V rows -> caches
            ph, pl, _ = eng._ptab_pair(L)
            eng._k("attention", "masr_relpos_attention_tc", _p(qkv), 3 * d, C, kvh.data_ptr(), kvl.data_ptr(), kvh.data_ptr() + 2 * d,
                   kvl.data_ptr() + 2 * d, 2 * d, cap, _p(ph), _p(pl), d, _p(L.pos_u), _p(L.pos_v), None, _p(t1p[0]), _p(t1p[1]), d, C,
                   qlen, klen, S, eng.h, eng.dk, C)
            # conv module over [cache ++ chunk] per slot (convolution.py:101-109): t0 <- norm_conv(x + out_proj(att))
            if fz:
                eng._tc_ln(t1p, d, tw[i, "wo"], L.bo, M, d, 1.0, x, L.ln_conv, t0p, y2=t0)
            else:
                eng._tc(t1p, d, tw[i, "wo"], L.bo, M, d, d, EPI_RESIDUAL, 1.0, x, d, C=x, ldc=d)
                eng._ln(x, L.ln_conv, t0, M)
            xc = self.xcat[i]                                   # [S, 14+16, d]
            xc[:, self.lorder:].copy_(t0.view(S, C, d))
            eng._k("affine_split", "masr_affine_split_f16", _p(xc), None, None, _p(xcp[0]), _p(xcp[1]), S * LC, d)
            eng._tc(xcp, d, tw[i, "pw1"], L.pw1_b, S * LC, 2 * d, d, EPI_BIAS_GLU, C=g, ldc=d)
            eng._k("dwconv_ln_silu", "masr_dwconv_ln_silu_f32", _p(g), d, LC, _p(L.dw), _p(L.dw_b), _p(L.cn[0]), _p(L.cn[1]), None, None,
                   _p(t1p[0]), _p(t1p[1]), d, C, _p(b["clen"]), S, d, w.kernel, 0, C, 1e-5)
            # new left context = the last `lorder` VALID rows: rows [n, n+lorder) of [cache ++ chunk], n = valid chunk rows
            self._shift(xc, None, LC, self.lorder, d * 4, self.QLEN)
            if fz:
                eng._tc_ln(t1p, d, tw[i, "pw2"], L.pw2_b, M, d, 1.0, x, L.ln_ff, t0p)
            else:
                eng._tc(t1p, d, tw[i, "pw2"], L.pw2_b, M, d, d, EPI_RESIDUAL, 1.0, x, d, C=x, ldc=d)
                eng._ln_split(x, L.ln_ff, t0p, M)
            eng._tc(t0p, d, tw[i, "ff1"], L.ff[1], M, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            if fz:      # x <- norm_final(x + 0.5 ffn), t0p <- the next block's norm_ff_macaron (or after_norm) of it
                eng._tc_ln(hidp, w.ffn, tw[i, "ff2"], L.ff[3], M, w.ffn, 0.5, x, L.ln_final, t0p,
                           ln2=w.layers[i + 1].ln_ffm if i + 1 < nl else w.after_norm)
            else:
                eng._tc(hidp, w.ffn, tw[i, "ff2"], L.ff[3], M, d, w.ffn, EPI_RESIDUAL, 0.5, x, d, C=x, ldc=d)
                eng._ln(x, L.ln_final, x, M)
        if not fz:
            eng._ln_split(x, w.after_norm, t0p, M)
        eng._tc(t0p, d, tw["ctc"], w.ctc_b, M, eng.V, d, C=b["logits"], ldc=eng.Vpad)
        eng._k("ctc_argmax", "masr_ctc_frame_argmax_f32", _p(b["logits"]), eng.Vpad, M, eng.V, _p(b["ids"]), _p(b["maxp"]), _p(self.probs), eng.V)


class SqueezeformerStreamPool(_PoolBase):
    """``ConformerStreamPool`` for the streaming Squeezeformer (``SqueezeformerEncoder.forward_chunk``,
    masr/model_utils/squeezeformer/encoder.py:240-361, with ``required_cache_size < 0`` as ``predict_stream`` passes).

    Blocks 5..10 run at half the frame rate: 16 chunk frames -> 8 (``TimeReductionLayerStream``: kernel 1, stride 2), and
    their K|V caches are kept at that rate (the reference stores them ``repeat_interleave``d to the full rate and reads them
    back with ``[::2]``, :339-356 — a round trip).  The conv-module left context (30 rows of the ada-scaled input,
    convolution.py:119-127) is kept as the fp16 (h,l) operand pair the pointwise GEMM consumes.  A chunk shorter than 67
    frames (the last one of a stream) is supported once per stream: afterwards the slot must be reset."""

    SHORT_ONCE = True

    def __init__(self, eng, n_slots: int, max_frames: int = 3000, use_graph: bool = True, keep_probs: bool = False):
        if not eng.causal:
            raise Exception("chunk decoding needs a streaming (causal) model")
        self._init_common(eng, n_slots, use_graph, keep_probs)
        self.cap = (max_frames + 15) // 16 * 16
        self.cap2 = self.cap // 2
        dev, d, w = eng.device, eng.d, eng.w
        f16, f32 = torch.float16, torch.float32
        nl = len(w.layers)
        S, C = n_slots, CHUNK_OUT
        C2 = C // 2
        self.lorder = w.kernel - 1
        F1 = (CHUNK_FRAMES - 1) // 2
        TH = (F1 + 1) // 2
        M = S * C
        LC = self.lorder + C
        self.reduced = [eng.REDUCE <= i < eng.RECOVER for i in range(nl)]
        self.kv = [(torch.zeros(S * (self.cap2 if r else self.cap), 2 * d, device=dev, dtype=f16),
                    torch.zeros(S * (self.cap2 if r else self.cap), 2 * d, device=dev, dtype=f16)) for r in self.reduced]
        # [cache ++ chunk] input rows of every block's conv module, as fp16 pairs
        self.xcat = [(torch.zeros(S, self.lorder + (C2 if r else C), d, device=dev, dtype=f16),
                      torch.zeros(S, self.lorder + (C2 if r else C), d, device=dev, dtype=f16)) for r in self.reduced]
        self.b = {
            "c1p": (torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16), torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16)),
            "c2p": (torch.empty(M * eng.f2, d, device=dev, dtype=f16), torch.empty(M * eng.f2, d, device=dev, dtype=f16)),
            "x": torch.zeros(M, d, device=dev, dtype=f32), "y": torch.zeros(M, d, device=dev, dtype=f32),
            "saved": torch.zeros(M, d, device=dev, dtype=f32),
            "t0p": (torch.zeros(M, d, device=dev, dtype=f16), torch.zeros(M, d, device=dev, dtype=f16)),
            "t1p": (torch.zeros(M, d, device=dev, dtype=f16), torch.zeros(M, d, device=dev, dtype=f16)),
            "hidp": (torch.empty(M, w.ffn, device=dev, dtype=f16), torch.empty(M, w.ffn, device=dev, dtype=f16)),
            "qkv": torch.zeros(M, 3 * d, device=dev, dtype=f32),
            "qkvp": (torch.zeros(M, 3 * d, device=dev, dtype=f16), torch.zeros(M, 3 * d, device=dev, dtype=f16)),
            "g": torch.empty(S * LC, d, device=dev, dtype=f32),
            "logits": torch.empty(M, eng.Vpad, device=dev, dtype=f32),
            "ids": torch.empty(M, device=dev, dtype=torch.int32), "maxp": torch.empty(M, device=dev, dtype=f32),
            "clen": torch.full((S,), self.lorder + C, device=dev, dtype=torch.int32),
            "clen2": torch.full((S,), self.lorder + C2, device=dev, dtype=torch.int32),
        }

    def reset(self, slot: int):
        """``InferencePredictor.reset_stream`` for one slot (inference_predictor.py:97-102)."""
        self.lens_host[slot] = 0
        for xh, xl in self.xcat:
            xh[slot].zero_()
            xl[slot].zero_()

    def _body(self):
        eng, S, C = self.eng, self.S, CHUNK_OUT
        C2 = C // 2
        w, d, tw, b = eng.w, eng.d, eng._tcw, self.b
        feats = self.feats_in
        M, M2 = S * C, S * C2
        F1 = (CHUNK_FRAMES - 1) // 2
        x, y, saved, t0p, t1p, hidp, qkv, qkvp, g = (b["x"], b["y"], b["saved"], b["t0p"], b["t1p"], b["hidp"], b["qkv"],
                                                      b["qkvp"], b["g"])
        eng._k("conv1", "masr_conv1_cmvn_relu_planes_f16", _p(feats), _p(w.cmvn_mean), _p(w.cmvn_istd), _p(w.conv1_w), _p(w.conv1_b),
               _p(b["c1p"][0]), _p(b["c1p"][1]), S, CHUNK_FRAMES, w.idim, F1, eng.w1_cols, d)
        eng._k("conv2", "masr_conv2_tc_f16x2", _p(b["c1p"][0]), _p(b["c1p"][1]), _p(tw["conv2"][0]), _p(tw["conv2"][1]), _p(w.conv2_b),
               None, _p(b["c2p"][0]), _p(b["c2p"][1]), S, F1, C, d)
        eng._tc(b["c2p"], eng.f2 * d, tw["embed"], w.embed_b, M, d, eng.f2 * d, EPI_BIAS, C=y, ldc=d)
        eng._ln_ada(y, w.preln, x, w.layers[0].att_ada, t0p, M)
        nl = len(w.layers)
        full = (M, C, self.QLEN, self.KLEN, self.BASE, b["clen"], self.cap)
        half = (M2, C2, self.QLEN2, self.KLEN2, self.BASE2, b["clen2"], self.cap2)
        Mi, Ci, rq, rk, rb, clen, cap = full
        for i, L in enumerate(w.layers):
            if i == eng.REDUCE:
                saved.copy_(x)
                eng._k("time_reduce", "masr_time_reduce_dw_split_f16", _p(x), C, _p(w.tr_dw), _p(w.tr_dw_b), _p(t1p[0]), _p(t1p[1]),
                       C2, self._m(self.QLEN), S, C2, int(w.tr_dw.shape[1]), 0, d)
                eng._tc(t1p, d, tw["tr_pw"], w.tr_pw_b, M2, d, d, EPI_BIAS, C=x, ldc=d)
                eng._k("affine_split", "masr_affine_split_f16", _p(x), _p(L.att_ada[0]), _p(L.att_ada[1]), _p(t0p[0]), _p(t0p[1]), M2, d)
                Mi, Ci, rq, rk, rb, clen, cap = half
            if i == eng.RECOVER:
                eng._k("affine_split", "masr_affine_split_f16", _p(x), None, None, _p(t1p[0]), _p(t1p[1]), M2, d)
                eng._tc(t1p, d, tw["rec"], w.rec_b, M2, d, d, EPI_BIAS, C=y, ldc=d)
                eng._k("upsample_add", "masr_upsample2_add_f32", _p(saved), _p(y), _p(x), C, C2, S, C, d)
                Mi, Ci, rq, rk, rb, clen, cap = full
                eng._k("affine_split", "masr_affine_split_f16", _p(x), _p(L.att_ada[0]), _p(L.att_ada[1]), _p(t0p[0]), _p(t0p[1]), Mi, d)
            LCi = self.lorder + Ci
            # MHA over [cache ++ chunk] keys
            eng._tc(t0p, d, tw[i, "qkv"], L.bqkv, Mi, 3 * d, d, C=qkv, Cp=qkvp, ldc=3 * d)
            kvh, kvl = self.kv[i]
            self._append_pair(qkvp, 3 * d, d, 2 * d, (kvh, kvl), cap, rb, rq, Ci)
            ph, pl, _ = eng._ptab_pair(L)
            eng._k("attention", "masr_relpos_attention_tc", _p(qkv), 3 * d, Ci, kvh.data_ptr(), kvl.data_ptr(), kvh.data_ptr() + 2 * d,
                   kvl.data_ptr() + 2 * d, 2 * d, cap, _p(ph), _p(pl), d, _p(L.pos_u), _p(L.pos_v), None, _p(t1p[0]), _p(t1p[1]), d, Ci,
                   self._m(rq), self._m(rk), S, eng.h, eng.dk, Ci)
            if self.fuse_ln:
                eng._tc_postln(t1p, d, tw[i, "wo"], L.bo, Mi, d, x, L.ln1, L.ffn1_ada, t0p)
            else:
                eng._tc(t1p, d, tw[i, "wo"], L.bo, Mi, d, d, EPI_RESIDUAL, 1.0, x, d, C=y, ldc=d)
                eng._ln_ada(y, L.ln1, x, L.ffn1_ada, t0p, Mi)
            eng._tc(t0p, d, tw[i, "f1a"], L.ffn1[1], Mi, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            if self.fuse_ln:
                eng._tc_postln(hidp, w.ffn, tw[i, "f1b"], L.ffn1[3], Mi, w.ffn, x, L.ln2, L.conv_ada, t0p)
            else:
                eng._tc(hidp, w.ffn, tw[i, "f1b"], L.ffn1[3], Mi, d, w.ffn, EPI_RESIDUAL, 1.0, x, d, C=y, ldc=d)
                eng._ln_ada(y, L.ln2, x, L.conv_ada, t0p, Mi)
            # conv module over [cache ++ chunk] per slot
            xh, xl = self.xcat[i]
            xh[:, self.lorder:].copy_(t0p[0][:Mi].view(S, Ci, d))
            xl[:, self.lorder:].copy_(t0p[1][:Mi].view(S, Ci, d))
            eng._tc((xh, xl), d, tw[i, "pw1"], L.pw1_b, S * LCi, 2 * d, d, EPI_BIAS_GLU, C=g, ldc=d)
            eng._k("dwconv_bn_silu", "masr_dwconv_bn_silu_f32", _p(g), d, LCi, _p(L.dw), _p(L.dw_b), _p(L.bn[0]), _p(L.bn[1]), None,
                   None, _p(t1p[0]), _p(t1p[1]), d, Ci, _p(clen), S, d, L.kernel, 0, Ci)
            # new left context = the last `lorder` VALID rows: rows [n, n + lorder) of [cache ++ chunk], n = valid chunk rows
            self._shift(xh, xl, LCi, self.lorder, d * 2, rq)
            if self.fuse_ln:
                eng._tc_postln(t1p, d, tw[i, "pw2"], L.pw2_b, Mi, d, x, L.ln3, L.ffn2_ada, t0p)
            else:
                eng._tc(t1p, d, tw[i, "pw2"], L.pw2_b, Mi, d, d, EPI_RESIDUAL, 1.0, x, d, C=y, ldc=d)
                eng._ln_ada(y, L.ln3, x, L.ffn2_ada, t0p, Mi)
            eng._tc(t0p, d, tw[i, "f2a"], L.ffn2[1], Mi, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            nxt = w.layers[i + 1].att_ada if (i + 1 < nl and i + 1 not in (eng.REDUCE, eng.RECOVER)) else None
            if self.fuse_ln:                            # (last block: pair(x) feeds the CTC head)
                eng._tc_postln(hidp, w.ffn, tw[i, "f2b"], L.ffn2[3], Mi, w.ffn, x, L.ln4, nxt, t0p)
            else:
                eng._tc(hidp, w.ffn, tw[i, "f2b"], L.ffn2[3], Mi, d, w.ffn, EPI_RESIDUAL, 1.0, x, d, C=y, ldc=d)
                eng._ln_ada(y, L.ln4, x, nxt, t0p, Mi)
        eng._tc(t0p, d, tw["ctc"], w.ctc_b, M, eng.V, d, C=b["logits"], ldc=eng.Vpad)
        eng._k("ctc_argmax", "masr_ctc_frame_argmax_f32", _p(b["logits"]), eng.Vpad, M, eng.V, _p(b["ids"]), _p(b["maxp"]), _p(self.probs), eng.V)


class EfficientConformerStreamPool(_PoolBase):
    """Batched chunk decoding for the streaming EfficientConformer (``EfficientConformerEncoder.forward_chunk``,
    masr/model_utils/efficient_conformer/encoder.py:267-392, ``required_cache_size < 0``).

    Blocks 0-3 run at 40 ms frames with grouped attention over a float32 K|V cache (keys regrouped from key 0, queries from
    the first chunk frame, attention.py:44-60); block 3's conv module strides by 2 (16 -> 8 frames, AvgPool residual);
    blocks 4-11 run at 80 ms frames (depthwise kernel 7) over fp16-pair K|V caches kept at that rate — the reference stores
    them ``repeat_interleave``d and reads them back with ``[::2]`` (:344,372), a round trip.  The caller's offset counts
    output frames and is doubled inside (:306).  One chunk yields 8 output frames.  A short final chunk is supported once
    per stream (reset afterwards)."""

    SHORT_ONCE = True
    OUT_ROWS = CHUNK_OUT // 2

    def __init__(self, eng, n_slots: int, max_frames: int = 3000, use_graph: bool = True, keep_probs: bool = False):
        if not eng.causal:
            raise Exception("chunk decoding needs a streaming (causal) model")
        self._init_common(eng, n_slots, use_graph, keep_probs)
        self.cap = (max_frames + 15) // 16 * 16
        self.cap2 = self.cap // 2
        dev, d, w = eng.device, eng.d, eng.w
        f16, f32 = torch.float16, torch.float32
        S, C = n_slots, CHUNK_OUT
        C2 = C // 2
        self.SL = eng.STRIDE_LAYER
        F1 = (CHUNK_FRAMES - 1) // 2
        TH = (F1 + 1) // 2
        M = S * C
        self.kv32 = {i: torch.zeros(S * self.cap, 2 * d, device=dev, dtype=f32) for i, L in enumerate(w.layers) if L.grouped}
        self.kv = {i: (torch.zeros(S * (self.cap if i <= self.SL else self.cap2), 2 * d, device=dev, dtype=f16),
                       torch.zeros(S * (self.cap if i <= self.SL else self.cap2), 2 * d, device=dev, dtype=f16))
                   for i, L in enumerate(w.layers) if not L.grouped}
        self.xcat = [(torch.zeros(S, (L.kernel - 1) + (C if i <= self.SL else C2), d, device=dev, dtype=f16),
                      torch.zeros(S, (L.kernel - 1) + (C if i <= self.SL else C2), d, device=dev, dtype=f16))
                     for i, L in enumerate(w.layers)]
        LCmax = max(x[0].shape[1] for x in self.xcat)
        self.b = {
            "c1p": (torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16), torch.zeros(4 * S * TH * 20 * d, device=dev, dtype=f16)),
            "c2p": (torch.empty(M * eng.f2, d, device=dev, dtype=f16), torch.empty(M * eng.f2, d, device=dev, dtype=f16)),
            "x": torch.zeros(M, d, device=dev, dtype=f32), "t0": torch.zeros(M, d, device=dev, dtype=f32),
            "t0p": (torch.zeros(M, d, device=dev, dtype=f16), torch.zeros(M, d, device=dev, dtype=f16)),
            "t1p": (torch.zeros(M, d, device=dev, dtype=f16), torch.zeros(M, d, device=dev, dtype=f16)),
            "hidp": (torch.empty(M, w.ffn, device=dev, dtype=f16), torch.empty(M, w.ffn, device=dev, dtype=f16)),
            "qb": torch.zeros(M, d, device=dev, dtype=f32), "kvn": torch.zeros(M, 2 * d, device=dev, dtype=f32),
            "qkv": torch.zeros(M, 3 * d, device=dev, dtype=f32),
            "qkvp": (torch.zeros(M, 3 * d, device=dev, dtype=f16), torch.zeros(M, 3 * d, device=dev, dtype=f16)),
            "g": torch.empty(S * LCmax, d, device=dev, dtype=f32),
            "logits": torch.empty(S * C2, eng.Vpad, device=dev, dtype=f32),
            "ids": torch.empty(S * C2, device=dev, dtype=torch.int32), "maxp": torch.empty(S * C2, device=dev, dtype=f32),
            "clen": torch.full((S,), LCmax, device=dev, dtype=torch.int32),
        }

    def reset(self, slot: int):
        self.lens_host[slot] = 0
        for xh, xl in self.xcat:
            xh[slot].zero_()
            xl[slot].zero_()

    def _body(self):
        eng, S, C = self.eng, self.S, CHUNK_OUT
        C2 = C // 2
        w, d, tw, b = eng.w, eng.d, eng._tcw, self.b
        feats = self.feats_in
        M, M2 = S * C, S * C2
        F1 = (CHUNK_FRAMES - 1) // 2
        x, t0, t0p, t1p, hidp, qkv, qkvp, g, qb, kvn = (b["x"], b["t0"], b["t0p"], b["t1p"], b["hidp"], b["qkv"], b["qkvp"],
                                                         b["g"], b["qb"], b["kvn"])
        eng._k("conv1", "masr_conv1_cmvn_relu_planes_f16", _p(feats), _p(w.cmvn_mean), _p(w.cmvn_istd), _p(w.conv1_w), _p(w.conv1_b),
               _p(b["c1p"][0]), _p(b["c1p"][1]), S, CHUNK_FRAMES, w.idim, F1, eng.w1_cols, d)
        eng._k("conv2", "masr_conv2_tc_f16x2", _p(b["c1p"][0]), _p(b["c1p"][1]), _p(tw["conv2"][0]), _p(tw["conv2"][1]), _p(w.conv2_b),
               None, _p(b["c2p"][0]), _p(b["c2p"][1]), S, F1, C, d)
        eng._tc(b["c2p"], eng.f2 * d, tw["embed"], w.embed_b, M, d, eng.f2 * d, EPI_BIAS_SCALE, float(d) ** 0.5, C=x, ldc=d)
        for i, L in enumerate(w.layers):
            half = i > self.SL
            Mi, Ci = (M2, C2) if half else (M, C)
            rq, rk, rb = (self.QLEN2, self.KLEN2, self.BASE2) if half else (self.QLEN, self.KLEN, self.BASE)
            cap = self.cap2 if half else self.cap
            lorder = L.kernel - 1
            LCi = lorder + Ci
            eng._ln_split(x, L.ln_ffm, t0p, Mi)
            eng._tc(t0p, d, tw[i, "ffm1"], L.ffm[1], Mi, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            eng._tc(hidp, w.ffn, tw[i, "ffm2"], L.ffm[3], Mi, d, w.ffn, EPI_RESIDUAL, 0.5, x, d, C=x, ldc=d)
            eng._ln_split(x, L.ln_mha, t0p, Mi)
            if L.grouped:
                wh, wl = tw[i, "qkv"]
                eng._tc(t0p, d, (wh[:d], wl[:d]), L.bqkv[:d], Mi, d, d, C=qb, ldc=d)
                eng._tc(t0p, d, (wh[d:], wl[d:]), L.bqkv[d:], Mi, 2 * d, d, C=kvn, ldc=2 * d)
                kc = self.kv32[i]
                self._append_pair(kvn, 2 * d, 0, 2 * d, kc, cap, rb, rq, Ci, elem_bytes=4)
                eng._k("attention", "masr_grouped_attention_cache_f32", _p(qb), d, Ci, kc.data_ptr(), kc.data_ptr() + 4 * d, 2 * d, cap,
                       _p(L.ptab), _p(L.pos_u), _p(L.pos_v), None, _p(t1p[0]), _p(t1p[1]), self._m(rq), self._m(rk), S, eng.h, eng.dk,
                       eng.GROUP, Ci)
            else:
                eng._tc(t0p, d, tw[i, "qkv"], L.bqkv, Mi, 3 * d, d, C=qkv, Cp=qkvp, ldc=3 * d)
                kvh, kvl = self.kv[i]
                self._append_pair(qkvp, 3 * d, d, 2 * d, (kvh, kvl), cap, rb, rq, Ci)
                ph, pl, _ = eng._ptab_pair(L)
                eng._k("attention", "masr_relpos_attention_tc", _p(qkv), 3 * d, Ci, kvh.data_ptr(), kvl.data_ptr(), kvh.data_ptr() + 2 * d,
                       kvl.data_ptr() + 2 * d, 2 * d, cap, _p(ph), _p(pl), d, _p(L.pos_u), _p(L.pos_v), None, _p(t1p[0]), _p(t1p[1]), d, Ci,
                       self._m(rq), self._m(rk), S, eng.h, eng.dk, Ci)
            eng._tc(t1p, d, tw[i, "wo"], L.bo, Mi, d, d, EPI_RESIDUAL, 1.0, x, d, C=x, ldc=d)
            # conv module over [cache ++ chunk] per slot (convolution.py:93-111)
            eng._ln_split(x, L.ln_conv, t0p, Mi)
            xh, xl = self.xcat[i]
            xh[:, lorder:].copy_(t0p[0][:Mi].view(S, Ci, d))
            xl[:, lorder:].copy_(t0p[1][:Mi].view(S, Ci, d))
            eng._tc((xh, xl), d, tw[i, "pw1"], L.pw1_b, S * LCi, 2 * d, d, EPI_BIAS_GLU, C=g, ldc=d)
            if i == self.SL:
                eng._k("dwconv_ln_silu", "masr_dwconv_ln_silu_strided_f32", _p(g), d, LCi, _p(L.dw), _p(L.dw_b), _p(L.cn[0]), _p(L.cn[1]),
                       None, None, _p(t1p[0]), _p(t1p[1]), d, C2, _p(b["clen"]), S, d, L.kernel, 0, 2, C2, 1e-5)
                eng._k("avgpool", "masr_avgpool2_time_f32", _p(x), C, _p(t0), C2, self._m(self.QLEN), S, C2, d)
                Mo, res = M2, t0
            else:
                eng._k("dwconv_ln_silu", "masr_dwconv_ln_silu_f32", _p(g), d, LCi, _p(L.dw), _p(L.dw_b), _p(L.cn[0]), _p(L.cn[1]), None,
                       None, _p(t1p[0]), _p(t1p[1]), d, Ci, _p(b["clen"]), S, d, L.kernel, 0, Ci, 1e-5)
                Mo, res = Mi, x
            # new left context = the last `lorder` VALID rows of [cache ++ chunk]
            self._shift(xh, xl, LCi, lorder, d * 2, rq)
            eng._tc(t1p, d, tw[i, "pw2"], L.pw2_b, Mo, d, d, EPI_RESIDUAL, 1.0, res, d, C=x, ldc=d)
            eng._ln_split(x, L.ln_ff, t0p, Mo)
            eng._tc(t0p, d, tw[i, "ff1"], L.ff[1], Mo, w.ffn, d, EPI_BIAS_SILU, Cp=hidp, ldc=w.ffn)
            eng._tc(hidp, w.ffn, tw[i, "ff2"], L.ff[3], Mo, d, w.ffn, EPI_RESIDUAL, 0.5, x, d, C=x, ldc=d)
            eng._ln(x, L.ln_final, x, Mo)
        eng._ln_split(x, w.after_norm, t0p, M2)
        eng._tc(t0p, d, tw["ctc"], w.ctc_b, M2, eng.V, d, C=b["logits"], ldc=eng.Vpad)
        eng._k("ctc_argmax", "masr_ctc_frame_argmax_f32", _p(b["logits"]), eng.Vpad, M2, eng.V, _p(b["ids"]), _p(b["maxp"]), _p(self.probs), eng.V)


class PoolStream:
    """One stream = a one-slot pool behind the single-stream interface ``MASRPredictor.predict_stream`` uses
    (``eng.new_stream()`` / ``eng.encode_chunk(chunk, stream)``)."""

    def __init__(self, pool):
        self.pool = pool
        self.batch = torch.zeros(1, CHUNK_FRAMES, 80, device=pool.eng.device, dtype=torch.float32)

    def reset(self):
        self.pool.reset(0)

    def encode_chunk(self, feats_chunk: torch.Tensor, required_cache_size: int = -1):
        if required_cache_size >= 0:
            raise NotImplementedError("bounded attention caches (required_cache_size >= 0) are not implemented for this model; "
                                      "predict_stream always asks for the whole history")
        n = int(feats_chunk.shape[0])
        if n > CHUNK_FRAMES:
            raise ValueError(f"a chunk has at most {CHUNK_FRAMES} feature frames")
        if subsampled_len(n) == 0:
            return None
        self.batch[0, :n].copy_(feats_chunk)
        ids, maxp, tout = self.pool.step(self.batch, [n])
        probs = None if self.pool.probs is None else self.pool.probs[:tout[0]]
        self.last_logits = self.pool.b["logits"][:tout[0]]      # (the streaming beam search reads the chunk's logits)
        return ids[0, :tout[0]], maxp[0, :tout[0]], probs


def make_pool(eng, n_slots: int, max_frames: int = 3000):
    """The batched chunk-decoding pool that matches the engine's model family."""
    from .engine import EfficientConformerEngine
    from .squeezeformer import SqueezeformerEngine
    if isinstance(eng, SqueezeformerEngine):
        return SqueezeformerStreamPool(eng, n_slots, max_frames)
    if isinstance(eng, EfficientConformerEngine):
        return EfficientConformerStreamPool(eng, n_slots, max_frames)
    if type(eng) is ConformerEngine:
        return ConformerStreamPool(eng, n_slots, max_frames)
    raise NotImplementedError(f"no stream pool for {type(eng).__name__}")


class StreamPool:
    """`predict_stream` for many streams at once (same per-stream results as one ``MASRPredictor`` per stream).

    Host work per push is O(slots) integer bookkeeping: the un-consumed feature frames of every slot live in ONE device
    ring ``[S, RING, 80]`` (appended and windowed with two batched index ops per push / per round instead of per-slot
    tensor ops), and the greedy history is folded incrementally — the reference re-collapses the whole history on every
    chunk (ctc_greedy_decoder.py:81-88), which yields the same tokens and the same left-to-right float32 score sum."""

    RING = 1024                    # feature frames kept per slot (un-consumed frames never exceed one push + one window)

    def __init__(self, eng: ConformerEngine, vocab: Sequence[str], n_slots: int, use_db_normalization: bool = True,
                 target_db: float = -20.0, max_frames: int = 3000):
        self.eng, self.vocab, self.S = eng, list(vocab), n_slots
        self.pool = make_pool(eng, n_slots, max_frames)
        self.use_db, self.target_db = use_db_normalization, target_db
        self.remained: List[Optional[np.ndarray]] = [None] * n_slots
        dev = eng.device
        # row S*RING is an all-zero frame: the source of window rows beyond a slot's chunk
        self.ring = torch.zeros(n_slots * self.RING + 1, 80, device=dev, dtype=torch.float32)
        self.head = [0] * n_slots      # absolute index of the first un-consumed frame
        self.count = [0] * n_slots     # un-consumed frames in the ring
        self._reset_hist(range(n_slots))

    def _reset_hist(self, slots):
        if not hasattr(self, "toks"):
            self.toks = [[] for _ in range(self.S)]
            self.prev = np.full(self.S, -1, np.int64)          # last frame id per slot (-1: none yet)
            self.acc = np.zeros(self.S, np.float32)            # left-to-right float32 sum of the non-blank max-probabilities
            self.nprob = np.zeros(self.S, np.int64)
        for s in slots:
            self.toks[s] = []
            self.prev[s], self.acc[s], self.nprob[s] = -1, np.float32(0.0), 0

    def _fold(self, ids_h: np.ndarray, mp_h: np.ndarray, tout: Sequence[int]):
        """Incremental ``greedy_decoder_chunk`` (ctc_greedy_decoder.py:70-89) for every slot at once: collapse repeats against
        the previous frame, drop blanks, and keep the score as the reference's left-to-right float32 sum — one vectorised
        float32 add per frame column (16 columns), so each slot's sum sees its terms in order with float32 rounding at
        every step, exactly like the scalar loop it replaces (which cost ~0.85 ms per push of 64 streams)."""
        S, C = ids_h.shape
        tout = np.asarray(tout, np.int64)
        if not tout.any():
            return
        ids = ids_h.astype(np.int64)
        valid = np.arange(C)[None, :] < tout[:, None]
        prev_col = np.concatenate([self.prev[:, None], ids[:, :-1]], axis=1)
        nonblank = valid & (ids != 0)
        new_tok = nonblank & (ids != prev_col)
        acc = self.acc
        for t in range(C):
            col = nonblank[:, t]
            if col.any():
                acc = np.where(col, (acc + mp_h[:, t]).astype(np.float32), acc)
        self.acc = acc.astype(np.float32)
        self.nprob += nonblank.sum(1)
        for s in np.nonzero(new_tok.any(1))[0]:
            self.toks[s].extend(ids[s, new_tok[s]].tolist())
        has = tout > 0
        last = np.take_along_axis(ids, np.maximum(tout - 1, 0)[:, None], axis=1)[:, 0]
        self.prev = np.where(has, last, self.prev)

    def reset_stream(self, slot: int):
        self.pool.reset(slot)
        self.remained[slot] = None
        self.head[slot], self.count[slot] = 0, 0
        self._reset_hist([slot])

    def _dev_index(self, idx: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(idx).to(self.eng.device, non_blocking=False)

    def _grow_ring(self, need: int):
        """Re-allocate the feature ring with room for `need` un-consumed frames per slot (a single message longer than the
        ring — about 10 s — is legal: the reference's `predict_stream` accepts any message length)."""
        R0, S = self.RING, self.S
        R1 = R0
        while R1 < need:
            R1 *= 2
        ring = torch.zeros(S * R1 + 1, 80, device=self.eng.device, dtype=torch.float32)
        src, dst = [], []
        for s in range(S):
            if self.count[s]:
                f = np.arange(self.count[s], dtype=np.int64)
                src.append(s * R0 + (self.head[s] + f) % R0)
                dst.append(s * R1 + f)
            self.head[s] = 0
        if src:
            ring.index_copy_(0, self._dev_index(np.concatenate(dst)), self.ring.index_select(0, self._dev_index(np.concatenate(src))))
        self.ring, self.RING = ring, R1

    def push(self, audio: Dict[int, object], is_end: bool = False, channels: int = 1, samp_width: int = 2,
             on_error: str = "raise"):
        """audio: slot -> np.ndarray | PCM bytes (one push per slot).  -> slot -> {'text','score'} | None, exactly what
        ``MASRPredictor.predict_stream(chunk, is_end)`` would return for that stream.

        Per-slot isolation: every slot is validated (decodable input, gain within 300 dB, pool / position-table capacity,
        no chunk after a short final chunk) BEFORE any state changes; a slot that fails keeps its previous state, is left
        out of the batched rounds and its exception lands in ``self.last_errors[slot]`` — the other slots of the push are
        decoded normally, as the reference only fails the offending connection (infer_server.py:130-137).
        ``on_error="raise"`` then raises ``StreamSlotError`` (carrying ``errors`` and the healthy slots' ``results``);
        ``on_error="return"`` returns the healthy slots' results only."""
        eng, S = self.eng, self.S
        self.last_errors = {}
        errors = self.last_errors
        cand = {}
        for s in sorted(audio):
            a = audio[s]
            try:
                new = samples_to_float32(a) if isinstance(a, np.ndarray) else pcm_bytes_to_float32(a, channels, samp_width)
            except Exception as e:                    # undecodable message: only this slot fails
                errors[s] = e
                continue
            cand[s] = new if self.remained[s] is None else np.concatenate([self.remained[s], new])
        slots = sorted(cand)
        out: Dict[int, Optional[dict]] = {}
        if slots:
            # one batched fbank over every slot's un-consumed samples; the tail keeps the gain (predict.py:274)
            feats, frames, status = eng.fbank([cand[s] for s in slots], self.use_db, self.target_db)
            gains = eng.last_gain.cpu().numpy() if self.use_db else np.ones(len(slots), np.float32)
            status_h = status.cpu().numpy()
            Fmax = feats.shape[1]
            pool = self.pool
            cap = getattr(pool, "cap", None)
            max_len = getattr(getattr(eng, "w", None), "max_len", None)
            lens_host = getattr(pool, "lens_host", None)
            short_once = bool(getattr(pool, "SHORT_ONCE", False))
            good, pending = [], {}
            for j, s in enumerate(slots):
                if status_h[j] != 0:
                    errors[s] = ValueError("无法将段规范化到目标dB，音频增益已经超过max_gain_db (300.0dB)")
                    continue
                total = self.count[s] + frames[j]
                starts = chunk_starts(total, is_end)
                if lens_host is not None and starts:
                    new_out = sum(subsampled_len(min(c + CHUNK_FRAMES, total) - c) for c in starts)
                    have = lens_host[s]
                    if short_once and have % CHUNK_OUT and new_out:
                        errors[s] = AssertionError(f"stream slot {s}: a short (final) chunk was already decoded; reset the stream first")
                        continue
                    if (cap is not None and have + new_out > cap) or (max_len is not None and have + new_out >= max_len):
                        errors[s] = AssertionError("offset: {} + x.shape[1]: {} is larger than the max_len: {}".format(
                            have, new_out, min(v for v in (cap, max_len) if v is not None)))
                        continue
                good.append((j, s))
                pending[s] = starts
            need = max((self.count[s] + frames[j] for j, s in good), default=0)
            if need > self.RING:
                self._grow_ring(need)
            R = self.RING
            # ---- commit: nothing below can fail for a validated slot ----
            for j, s in good:
                tail = cand[s][FRAME_SHIFT * frames[j]:]
                self.remained[s] = (tail * np.float32(gains[j])).astype(np.float32) if self.use_db else tail
            # append the new frames of every slot to the ring: one gather + one scatter, index lists built without a per-slot loop
            gj = np.asarray([j for j, _ in good], np.int64)
            gs = np.asarray([s for _, s in good], np.int64)
            nf = np.asarray([frames[j] for j, _ in good], np.int64)
            if nf.sum() > 0:
                seg = np.repeat(np.arange(len(good)), nf)
                off = np.arange(int(nf.sum()), dtype=np.int64) - np.repeat(np.cumsum(nf) - nf, nf)
                start = np.asarray([self.head[s] + self.count[s] for s in gs], np.int64)
                src = gj[seg] * Fmax + off
                dst = gs[seg] * R + (start[seg] + off) % R
                self.ring.index_copy_(0, self._dev_index(dst), feats.view(-1, 80).index_select(0, self._dev_index(src)))
            for j, s in good:
                self.count[s] += frames[j]
            live = [s for _, s in good]
            ends = {}
            rounds = max((len(v) for v in pending.values()), default=0)
            zero_row = S * R
            win = np.arange(CHUNK_FRAMES, dtype=np.int64)
            for r in range(rounds):
                nfr = [0] * S
                act = [s for s in live if r < len(pending[s])]
                idx = np.full((S, CHUNK_FRAMES), zero_row, np.int64)
                if act:
                    a = np.asarray(act, np.int64)
                    cur = np.asarray([pending[s][r] for s in act], np.int64)
                    cnt = np.asarray([self.count[s] for s in act], np.int64)
                    n = np.minimum(cur + CHUNK_FRAMES, cnt) - cur
                    hd = np.asarray([self.head[s] for s in act], np.int64)
                    rows = a[:, None] * R + (hd[:, None] + cur[:, None] + win[None, :]) % R
                    idx[a] = np.where(win[None, :] < n[:, None], rows, zero_row)
                    for s, c_, n_ in zip(act, cur, n):
                        nfr[s] = int(n_)
                        ends[s] = int(c_ + n_)
                batch = self.ring.index_select(0, self._dev_index(idx.reshape(-1))).view(S, CHUNK_FRAMES, 80)
                ids, maxp, tout = self.pool.step(batch, nfr)
                self._fold(ids.cpu().numpy(), maxp.cpu().numpy(), tout)
            for s in live:
                if not pending[s]:
                    out[s] = None
                    continue
                consumed = ends[s] - CACHED_FEATURE_NUM              # predict.py:330: keep the last 3 frames of the window
                self.head[s] = (self.head[s] + consumed) % R
                self.count[s] -= consumed
                out[s] = {"text": ids_to_text(self.toks[s], self.vocab), "score": greedy_score(self.acc[s], int(self.nprob[s]))}
        if errors and on_error == "raise":
            raise StreamSlotError(errors, out)
        return out


class StreamSlotError(Exception):
    """Some slots of a ``StreamPool.push`` failed: ``errors`` maps slot -> exception, ``results`` holds what the healthy
    slots returned (their state advanced normally)."""

    def __init__(self, errors: Dict[int, Exception], results: Dict[int, Optional[dict]]):
        self.errors, self.results = dict(errors), dict(results)
        first = next(iter(self.errors.values()))
        super().__init__(f"{len(self.errors)} stream slot(s) failed: slot {next(iter(self.errors))}: {first}")
