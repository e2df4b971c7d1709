"""DeepSpeech2 engine (configs/deepspeech2.yml; masr/model_utils/deepspeech2/{conv,encoder,model}.py):
CMVN -> Conv2d(1,32,3,2)+ReLU -> Conv2d(32,32,3,2)+ReLU -> 5 x [LSTM(1024) uni (streaming) / bi -> LayerNorm] -> CTC.

Input projections and the CTC head are tensor-core GEMMs (FP16x2 split); the first projection (K = 608) and the
recurrence run on the fp32 FMA pipe, one launch per time step (replayed as a CUDA graph).  Whole-utterance batches and
the chunked streaming path with carried (h, c) state (inference_predictor.py:66-78) are both implemented."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import EPI_BIAS, call
from .engine import ConformerEngine, _p, subsampled_len


@dataclass
class DS2Weights:
    d_model: int          # encoder output width (H * dirs)
    heads: int
    ffn: int
    kernel: int
    idim: int
    vocab: int
    max_len: int
    hidden: int = 1024
    dirs: int = 1
    cmvn_mean: torch.Tensor = None
    cmvn_istd: torch.Tensor = None
    conv1_w: torch.Tensor = None
    conv1_b: torch.Tensor = None
    conv2_w: torch.Tensor = None
    conv2_b: torch.Tensor = None
    layers: list = field(default_factory=list)      # unused (ConformerEngine plumbing)
    rnn: List[dict] = field(default_factory=list)   # per layer: {"wih": [dirs], "whh": [dirs], "bias": [dirs], "ln": (g, b)}
    ctc_w: torch.Tensor = None
    ctc_b: torch.Tensor = None
    pe: torch.Tensor = None


def pack_deepspeech2(sd: Dict[str, torch.Tensor], device) -> DS2Weights:
    dev = torch.device(device)

    def D(t):
        return t.contiguous().to(dev)

    idim = sd["encoder.global_cmvn.mean"].shape[0]
    C = sd["encoder.conv.conv.0.weight"].shape[0]
    from .weights import check_supported
    check_supported(sd, "deepspeech2")
    H = sd["encoder.rnns.0.rnn.weight_hh_l0"].shape[1]
    dirs = 2 if "encoder.rnns.0.rnn.weight_hh_l0_reverse" in sd else 1
    nl = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.rnns."))
    vocab = sd["decoder.ctc_lo.weight"].shape[0]
    w = DS2Weights(d_model=H * dirs, heads=1, ffn=0, kernel=0, idim=idim, vocab=vocab, max_len=0, hidden=H, dirs=dirs)
    w.cmvn_mean, w.cmvn_istd = D(sd["encoder.global_cmvn.mean"]), D(sd["encoder.global_cmvn.istd"])
    w.conv1_w, w.conv1_b = D(sd["encoder.conv.conv.0.weight"].reshape(C, 9)), D(sd["encoder.conv.conv.0.bias"])
    w.conv2_w = D(sd["encoder.conv.conv.2.weight"].permute(0, 2, 3, 1).reshape(C, 9 * C))
    w.conv2_b = D(sd["encoder.conv.conv.2.bias"])
    f2 = ((idim - 1) // 2 - 1) // 2
    for l in range(nl):
        p = f"encoder.rnns.{l}.rnn."
        ent = {"wih": [], "whh": [], "bias": []}
        for suf in ("", "_reverse")[:dirs]:
            wih = sd[p + "weight_ih_l0" + suf]
            if l == 0:   # conv output is channels-last here: permute the (c*19+f) input columns to (f*32+c)
                wih = wih.reshape(4 * H, C, f2).permute(0, 2, 1).reshape(4 * H, f2 * C)
            ent["wih"].append(D(wih))
            ent["whh"].append(D(sd[p + "weight_hh_l0" + suf]))
            ent["bias"].append(D(sd[p + "bias_ih_l0" + suf] + sd[p + "bias_hh_l0" + suf]))
        ent["ln"] = (D(sd[f"encoder.rnns.{l}.layer_norm.weight"]), D(sd[f"encoder.rnns.{l}.layer_norm.bias"]))
        w.rnn.append(ent)
    w.ctc_w, w.ctc_b = D(sd["decoder.ctc_lo.weight"]), D(sd["decoder.ctc_lo.bias"])
    return w


class DeepSpeech2Stream:
    """(h, c) of the 5 LSTM layers carried between chunks (inference_predictor.py:45-46,97-99)."""

    def __init__(self, eng: "DeepSpeech2Engine"):
        self.eng = eng
        H, nl = eng.H, len(eng.w.rnn)
        self.hT = torch.zeros(nl, 2, H, 32, device=eng.device, dtype=torch.float32)   # ping-pong, transposed [H][32]
        self.c = torch.zeros(nl, 1, H, device=eng.device, dtype=torch.float32)
        self.cur = [0] * nl

    def reset(self):
        self.hT.zero_()
        self.c.zero_()
        self.cur = [0] * len(self.cur)


class DeepSpeech2Engine(ConformerEngine):
    def __init__(self, weights_src, streaming: bool = True, device: str = "cuda", max_len: int = 5000, gemm: str = "tc",
                 use_graphs: bool = True):
        if gemm != "tc":
            raise ValueError("DeepSpeech2Engine implements the tensor-core projection path only")
        super().__init__(weights_src, streaming, device, max_len, gemm, use_graphs)
        self.H = self.w.hidden
        self.dirs = self.w.dirs
        # one persistent launch per layer and direction (masr_lstm_seq_f32) instead of one launch per time step
        self.persistent_lstm = os.environ.get("MASR_LSTM_PERSISTENT", "1") != "0"
        if bool(streaming) != (self.dirs == 1):
            raise Exception("streaming DeepSpeech2 needs forward-only LSTM weights, non-streaming bidirectional ones")

    def _pack(self, sd, max_len):
        return pack_deepspeech2(sd, self.device)

    def _precompute_pos(self):
        pass

    def _split_weights(self):
        t = self._tcw
        t["ctc"] = self._split(self.w.ctc_w)
        for l, ent in enumerate(self.w.rnn):
            if l > 0:
                t[l, "wih"] = [self._split(x) for x in ent["wih"]]
        torch.cuda.synchronize(self.device)

    def _workspace(self, B: int, Fmax: int):
        key = (B, Fmax)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev, f32, f16 = self.device, torch.float32, torch.float16
        F1 = (Fmax - 1) // 2
        T = subsampled_len(Fmax)
        M = max(1, B * T)
        C = self.w.conv1_w.shape[0]
        D = self.H * self.dirs
        nb = (B + 31) // 32
        ws = {
            "c1": torch.empty(B * max(1, F1) * self.w1_cols * C, device=dev, dtype=f32),
            "c2": torch.empty(M, self.f2 * C, device=dev, dtype=f32),
            "gx": torch.empty(M, 4 * self.H, device=dev, dtype=f32),
            "out": torch.zeros(M, D, device=dev, dtype=f32),
            "t0": torch.empty(M, D, device=dev, dtype=f32),
            "xp": (torch.empty(M, D, device=dev, dtype=f16), torch.empty(M, D, device=dev, dtype=f16)),
            "hT": torch.zeros(2, nb, self.H, 32, device=dev, dtype=f32),
            "c": torch.zeros(B, self.H, device=dev, dtype=f32),
            "logits": torch.empty(M, self.Vpad, device=dev, dtype=f32),
            "ids": torch.empty(M, device=dev, dtype=torch.int32),
            "maxp": torch.empty(M, device=dev, dtype=f32),
        }
        self._alloc_out_pack(ws, B, T)
        ws["t0p"] = ws["xp"]
        if len(self._ws) > 8:
            self._ws.clear()
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------------
    def _rnn_stack(self, ws, B, T, M, tlens, hT_init=None, c_init=None, stream: Optional[DeepSpeech2Stream] = None):
        """x = ws['c2'] [M, 608] -> ws['xp'] pair of the last LayerNorm output (and ws['t0'] fp32)."""
        w, H, dirs, D = self.w, self.H, self.dirs, self.H * self.dirs
        out, gx, xp = ws["out"], ws["gx"], ws["xp"]
        for l, ent in enumerate(w.rnn):
            for di in range(dirs):
                if l == 0:
                    K0 = ws["c2"].shape[1]
                    self._gemm(ws["c2"], K0, ent["wih"][di], ent["bias"][di], gx, 4 * H, M, 4 * H, K0, EPI_BIAS, tag="lstm_xproj")
                else:
                    self._tc(xp, D, self._tcw[l, "wih"][di], ent["bias"][di], M, 4 * H, D, EPI_BIAS, C=gx, ldc=4 * H, tag="lstm_xproj")
                if stream is None:
                    hT, c = ws["hT"], ws["c"]
                    hT.zero_()
                    c.zero_()
                    cur = 0
                else:
                    hT, c, cur = stream.hT[l].unsqueeze(1), stream.c[l], stream.cur[l]
                if self.persistent_lstm and H % 128 == 0 and H <= 1024:
                    # the whole recurrence of this layer / direction in one persistent launch (W_hh slices resident in shared memory)
                    if ws.get("lstm_ws") is None:
                        nbytes = _lib.C.c_int64(0)
                        call("masr_lstm_seq_workspace_bytes", B, H, _lib.C.byref(nbytes))
                        ws["lstm_ws"] = torch.empty(nbytes.value, device=self.device, dtype=torch.uint8)
                    self._k("lstm_seq", "masr_lstm_seq_f32", _p(gx), 4 * H, T, _p(ent["whh"][di]), _p(hT[cur]), _p(hT[1 - cur]), _p(c),
                            _p(out), None, None, D, di * H, _p(tlens), B, H, T, di, _p(ws["lstm_ws"]), ws["lstm_ws"].numel())
                    cur = 1 - cur
                else:
                    for s in range(T):
                        self._k("lstm_step", "masr_lstm_step_f32", _p(gx), 4 * H, T, _p(ent["whh"][di]), _p(hT[cur]), _p(hT[1 - cur]),
                                _p(c), _p(out), None, None, D, di * H, _p(tlens), B, H, s, di)
                        cur = 1 - cur
                if stream is not None:
                    stream.cur[l] = cur
            self._k("layernorm", "masr_layernorm_split_f16", _p(out), D, _p(ent["ln"][0]), _p(ent["ln"][1]), _p(xp[0]), _p(xp[1]),
                    D, M, D, 1e-5)
        last = w.rnn[-1]["ln"]
        self._k("layernorm", "masr_layernorm_f32", _p(out), D, _p(last[0]), _p(last[1]), _p(ws["t0"]), D, M, D, 1e-5)

    def _front(self, feats, ws, B, Fmax, F1, T):
        w = self.w
        C = w.conv1_w.shape[0]
        self._k("conv1", "masr_conv1_cmvn_relu_f32", _p(feats), _p(w.cmvn_mean), _p(w.cmvn_istd), _p(w.conv1_w), _p(w.conv1_b),
                _p(ws["c1"]), B, Fmax, w.idim, F1, self.w1_cols, C)
        self._k("conv2", "masr_conv2_s2_relu_f32", _p(ws["c1"]), _p(w.conv2_w), _p(w.conv2_b), _p(ws["c2"]), B, F1, self.w1_cols,
                T, self.f2, C)

    def encode(self, feats: torch.Tensor, feat_lens: Sequence[int], tlens_dev: Optional[torch.Tensor] = None):
        B, Fmax = feats.shape[0], feats.shape[1]
        F1 = (Fmax - 1) // 2
        T = subsampled_len(Fmax)
        tl = [subsampled_len(int(f)) for f in feat_lens]
        ws = self._workspace(B, Fmax)
        if T == 0:
            return ws["t0"][:0], tl, 0, ws
        M = B * T
        if tlens_dev is not None:
            ws["tlens"], ws["tl_host"] = tlens_dev, None
        elif ws.get("tl_host") != tl:
            ws["tlens"] = torch.tensor(tl, dtype=torch.int32, device=self.device)
            ws["tl_host"] = list(tl)
        self._front(feats, ws, B, Fmax, F1, T)
        self._rnn_stack(ws, B, T, M, ws["tlens"])
        return ws["t0"][:M], tl, T, ws

    def _ctc_operand(self, ws):
        return ws["xp"], self.H * self.dirs

    def ctc_logits(self, enc, ws):
        M = enc.shape[0]
        D = self.H * self.dirs
        self._tc(ws["xp"], D, self._tcw["ctc"], self.w.ctc_b, M, self.V, D, C=ws["logits"], ldc=self.Vpad, tag="ctc_head")
        return ws["logits"]

    # ---- streaming ----------------------------------------------------------------------------------
    def new_stream(self) -> DeepSpeech2Stream:
        if self.dirs != 1:
            raise Exception("chunk decoding needs a streaming (forward-only) model")
        return DeepSpeech2Stream(self)

    def encode_chunk(self, feats_chunk: torch.Tensor, st: DeepSpeech2Stream, required_cache_size: int = -1,
                     want_probs: bool = False):
        """``DeepSpeech2Model.get_encoder_out_chunk`` for one stream (model.py:70-77): feats [n, 80] on device ->
        (ids, max-prob[, posteriors]) for ((n-1)//2-1)//2 frames; the LSTM state is carried in ``st``."""
        n = int(feats_chunk.shape[0])
        T = subsampled_len(n)
        if T == 0:
            return None
        ws = self._workspace(1, n)
        if ws.get("tl_host") != [T]:
            ws["tlens"] = torch.tensor([T], dtype=torch.int32, device=self.device)
            ws["tl_host"] = [T]
        self._front(feats_chunk.reshape(1, n, -1), ws, 1, n, (n - 1) // 2, T)
        self._rnn_stack(ws, 1, T, T, ws["tlens"], stream=st)
        logits = self.ctc_logits(ws["t0"][:T], ws)
        probs = torch.empty(T, self.V, device=self.device, dtype=torch.float32) if want_probs else None
        self._k("ctc_argmax", "masr_ctc_frame_argmax_f32", _p(logits), self.Vpad, T, self.V, _p(ws["ids"]), _p(ws["maxp"]),
                _p(probs), self.V)
        st.last_logits = logits[:T]                    # (the streaming beam search reads the chunk's logits)
        return ws["ids"][:T], ws["maxp"][:T], probs
