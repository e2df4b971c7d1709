"""Oracle (test infrastructure): DeepSpeech2 inference forward, restated as plain torch-CPU functions
over a ``state_dict`` (B=1 semantics; whole utterance and chunked with carried LSTM state).

Follows masr/model_utils/deepspeech2/:
  * ``Conv2dSubsampling4Pure.forward``   conv.py:15-22   (32 channels, output [B, T, 32*19] in (c, f) order)
  * ``RNN.forward`` / ``CRNNEncoder``    encoder.py:36-45,96-129 (5 x [LSTM(1024) uni- or bi-directional -> LayerNorm])
  * ``DeepSpeech2Model.get_encoder_out[_chunk]``  model.py:65-77  (softmax(ctc_lo(.)) ; streaming -> 'forward' LSTM)
The LSTM cell is written out explicitly (PyTorch gate order i, f, g, o) rather than calling ``nn.LSTM``, so the
restatement is independent of cuDNN/oneDNN fused kernels; it is pinned against the reference in the golden tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class DS2Config:
    input_dim: int = 80
    layers: int = 5
    hidden: int = 1024
    bidirectional: bool = False       # streaming=True -> 'forward' (deepspeech2/model.py:42)


def subsample(sd, feats):
    x = (feats - sd["encoder.global_cmvn.mean"]) * sd["encoder.global_cmvn.istd"]
    x = x.unsqueeze(1)
    x = F.relu(F.conv2d(x, sd["encoder.conv.conv.0.weight"], sd["encoder.conv.conv.0.bias"], stride=2))
    x = F.relu(F.conv2d(x, sd["encoder.conv.conv.2.weight"], sd["encoder.conv.conv.2.bias"], stride=2))
    x = x.permute(0, 2, 1, 3)
    return x.reshape(x.shape[0], x.shape[1], -1)


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, h0, c0, reverse: bool):
    """x [T, in] -> (out [T, H], h_T, c_T); gates = W_ih x + b_ih + W_hh h + b_hh, order (i, f, g, o)."""
    T, H = x.shape[0], w_hh.shape[1]
    gx = F.linear(x, w_ih, b_ih)
    h, c = h0, c0
    out = x.new_zeros(T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = gx[t] + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g[:H], g[H:2 * H], g[2 * H:3 * H], g[3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out, h, c


def encode(sd, cfg: DS2Config, feats: torch.Tensor, state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """feats [1, F, 80] -> (enc [T, H or 2H], (h [L, dirs, H], c [L, dirs, H]))."""
    x = subsample(sd, feats)[0]
    H = cfg.hidden
    dirs = 2 if cfg.bidirectional else 1
    hs, cs = [], []
    for l in range(cfg.layers):
        p = f"encoder.rnns.{l}.rnn."
        outs, hl, cl = [], [], []
        for dname, rev in (("", False), ("_reverse", True))[:dirs]:
            di = 1 if rev else 0
            h0 = x.new_zeros(H) if state is None else state[0][l, di]
            c0 = x.new_zeros(H) if state is None else state[1][l, di]
            o, h, c = lstm_direction(x, sd[p + "weight_ih_l0" + dname], sd[p + "weight_hh_l0" + dname],
                                     sd[p + "bias_ih_l0" + dname], sd[p + "bias_hh_l0" + dname], h0, c0, rev)
            outs.append(o); hl.append(h); cl.append(c)
        x = torch.cat(outs, dim=1)
        x = F.layer_norm(x, (x.shape[1],), sd[f"encoder.rnns.{l}.layer_norm.weight"], sd[f"encoder.rnns.{l}.layer_norm.bias"], 1e-5)
        hs.append(torch.stack(hl)); cs.append(torch.stack(cl))
    return x, (torch.stack(hs), torch.stack(cs))


def get_encoder_out(sd, cfg, feats, state=None):
    """-> (probs [T, V], new state)."""
    enc, st = encode(sd, cfg, feats, state)
    return torch.softmax(F.linear(enc, sd["decoder.ctc_lo.weight"], sd["decoder.ctc_lo.bias"]), dim=1), st
