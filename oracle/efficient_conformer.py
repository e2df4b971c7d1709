"""Oracle (test infrastructure): EfficientConformer inference forward, restated as plain torch-CPU
functions over a ``state_dict`` (full-context pass; B=1 semantics).

Follows masr/model_utils/efficient_conformer/:
  * ``EfficientConformerEncoder.forward``            encoder.py:213-265 (grouped attention in blocks 0-3, strided conv
                                                     block 3, kernel 15 -> 7, pos_emb/masks re-strided :253-258)
  * ``StrideConformerEncoderLayer.forward``          encoder.py:454-545 (AvgPool1d(2,2,ceil_mode) on the residual :520-523)
  * ``GroupedRelPositionMultiHeadedAttention``       attention.py:35-69 (pad4group: the ``view`` regroups the [t,h,d_k]
                                                     memory of 3 consecutive frames into 4 heads x 192), :120-182
  * ``ConvolutionModule.forward`` (stride)           convolution.py:73-134
The un-grouped blocks and the FFN / subsampling are the Conformer ones (oracle/conformer.py).
Config trap: ``encoder_conf.efficient_conf`` is swallowed by ``**kwargs``; the constructor
defaults apply (stride_layer_idx=3, stride=2, group_layer_idx=(0,1,2,3), group_size=3, stride_kernel=True).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

from . import conformer as oc


@dataclass
class EfficientConfig(oc.ConformerConfig):
    stride_layer: int = 3
    stride: int = 2
    group_layers: tuple = (0, 1, 2, 3)
    group_size: int = 3

    def kernel_of(self, i: int) -> int:
        return self.kernel if i <= self.stride_layer else self.kernel // self.stride


def grouped_attention(sd, p, cfg: EfficientConfig, x, pos_emb):
    """x [B,T,d]; pos_emb [1,T,d] -> [B,T,d] (no cache, full mask)."""
    B, T, d = x.shape
    h, dk, g = cfg.heads, cfg.d_model // cfg.heads, cfg.group_size
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"])
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"])
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"])
    pos = F.linear(pos_emb, sd[p + "linear_pos.weight"])
    pad = (g - T % g) % g
    # [B,T,d] rows of h*dk floats; zero-pad time to a multiple of g and regroup 3 frames x 256 -> 4 heads x 192
    def regroup(t):
        t = F.pad(t, (0, 0, 0, pad))
        return t.reshape(t.shape[0], -1, h, dk * g).transpose(1, 2)          # [B,h,T/g,dk*g]
    qg, kg, vg, pg = regroup(q), regroup(k), regroup(v), regroup(pos)
    qu = qg + sd[p + "pos_bias_u"][None, :, None, :]
    qv = qg + sd[p + "pos_bias_v"][None, :, None, :]
    scores = (qu @ kg.transpose(-2, -1) + qv @ pg.transpose(-2, -1)) / math.sqrt(dk * g)
    ctx = torch.softmax(scores, dim=-1) @ vg                                 # [B,h,T/g,dk*g]
    ctx = ctx.transpose(1, 2).reshape(B, -1, d)
    if pad:
        ctx = ctx[:, :ctx.shape[1] - pad]
    return F.linear(ctx, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def conv_module(sd, p, cfg: EfficientConfig, x, kernel: int, stride: int):
    xt = x.transpose(1, 2)
    lorder = kernel - 1 if cfg.causal else 0
    if lorder > 0:
        xt = F.pad(xt, (lorder, 0))
    y = F.glu(F.conv1d(xt, sd[p + "pointwise_conv1.weight"], sd[p + "pointwise_conv1.bias"]), dim=1)
    y = F.conv1d(y, sd[p + "depthwise_conv.weight"], sd[p + "depthwise_conv.bias"], stride=stride,
                 padding=0 if lorder > 0 else (kernel - 1) // 2, groups=cfg.d_model)
    y = F.silu(F.layer_norm(y.transpose(1, 2), (cfg.d_model,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)).transpose(1, 2)
    y = F.conv1d(y, sd[p + "pointwise_conv2.weight"], sd[p + "pointwise_conv2.bias"])
    return y.transpose(1, 2)


def encoder_layer(sd, i, cfg: EfficientConfig, x, pos_emb):
    p = f"encoder.encoders.{i}."
    eps = cfg.ln_eps
    x = x + 0.5 * oc._ffn(sd, p + "feed_forward_macaron", oc._ln(sd, p + "norm_ff_macaron", x, eps))
    xn = oc._ln(sd, p + "norm_mha", x, eps)
    if i in cfg.group_layers:
        a = grouped_attention(sd, p + "self_attn.", cfg, xn, pos_emb)
    else:
        a, _ = oc.rel_attention(sd, p + "self_attn.", cfg, xn, pos_emb, None, None)
    x = x + a
    strided = i == cfg.stride_layer
    c = conv_module(sd, p + "conv_module.", cfg, oc._ln(sd, p + "norm_conv", x, eps), cfg.kernel_of(i),
                    cfg.stride if strided else 1)
    res = x
    if strided:
        res = F.avg_pool1d(x.transpose(1, 2), cfg.stride, cfg.stride, 0, ceil_mode=True, count_include_pad=False).transpose(1, 2)
    x = res + c
    x = x + 0.5 * oc._ffn(sd, p + "feed_forward", oc._ln(sd, p + "norm_ff", x, eps))
    return oc._ln(sd, p + "norm_final", x, eps)


def encode(sd, cfg: EfficientConfig, feats: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    x = oc.subsample(sd, cfg, feats)
    pos_emb = oc.sinusoid_table(cfg)[None, :x.shape[1]]
    for i in range(cfg.blocks):
        x = encoder_layer(sd, i, cfg, x, pos_emb)
        if i == cfg.stride_layer:
            pos_emb = pos_emb[:, ::cfg.stride]
        if taps is not None:
            taps[f"layer{i}"] = x.clone()
    return oc._ln(sd, "encoder.after_norm", x, cfg.ln_eps)


def get_encoder_out(sd, cfg, feats: torch.Tensor) -> torch.Tensor:
    return oc.ctc_probs(sd, encode(sd, cfg, feats))
