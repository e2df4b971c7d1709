"""Dev experiment (CPU, not a test): how many per-frame CTC argmaxes flip against the float32 oracle when the operands
of every dense contraction of the Conformer (linears, pointwise convs, conv-2, CTC head) are rounded the way a tensor-core
scheme would round them.  Guides the precision policy of csrc/tc_gemm.cu:

    a16w16   activations and weights rounded to fp16           (single-pass FP16 MMA)
    a16      activations fp16, weights kept to 22 bits (h + l)  (2 MMAs: Ah.Wh + Ah.Wl — no activation low part)
    w16      activations 22 bits, weights fp16                  (2 MMAs: Ah.Wh + Al.Wh)
    split    both 22 bits                                       (3 MMAs, the shipped scheme)

    python tests/precision_probe.py [n_utterances]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from masr_b200 import synth
from oracle import conformer as oc, fbank as ob

_linear, _conv1d, _conv2d = F.linear, F.conv1d, F.conv2d
MODE = {"a": None, "w": None}


def rnd(t, bits):
    if bits is None:
        return t
    if bits == 11:
        return t.half().float()
    h = t.half().float()                       # 22 bits: h + l, l = fp16((x - h) * 2^11)
    return h + ((t - h) * 2048.0).half().float() / 2048.0


def linear(x, w, b=None):
    return _linear(rnd(x, MODE["a"]), rnd(w, MODE["w"]), b)


def conv1d(x, w, b=None, *a, **k):
    if w.shape[-1] == 1 and k.get("groups", 1) == 1:          # pointwise convs only; the depthwise conv stays fp32
        return _conv1d(rnd(x, MODE["a"]), rnd(w, MODE["w"]), b, *a, **k)
    return _conv1d(x, w, b, *a, **k)


def conv2d(x, w, b=None, *a, **k):
    if w.shape[1] > 1:                                          # conv #2 (256 -> 256); conv #1 is an fp32 SIMT kernel
        return _conv2d(rnd(x, MODE["a"]), rnd(w, MODE["w"]), b, *a, **k)
    return _conv2d(x, w, b, *a, **k)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.set_num_threads(max(1, os.cpu_count() // 2))
    sd = synth.to_torch(synth.conformer_state_dict(0))
    cfg = oc.ConformerConfig()
    feats = [torch.from_numpy(ob.featurize(synth.noise_audio(1000 + i, 160000)))[None] for i in range(n)]
    F.linear, F.conv1d, F.conv2d = linear, conv1d, conv2d
    try:
        res = {}
        for name, (a, w) in {"fp32": (None, None), "split": (22, 22), "a16": (11, 22), "w16": (22, 11), "a16w16": (11, 11)}.items():
            MODE["a"], MODE["w"] = a, w
            ids = []
            with torch.no_grad():
                for f in feats:
                    ids.append(oc.get_encoder_out(sd, cfg, f)[0].argmax(1).numpy())
            res[name] = np.concatenate(ids)
        base = res["fp32"]
        print(f"{n} utterances x 10 s, {base.size} frames")
        for name, ids in res.items():
            print(f"{name:8s} flipped argmaxes: {int((ids != base).sum())}")
    finally:
        F.linear, F.conv1d, F.conv2d = _linear, _conv1d, _conv2d


if __name__ == "__main__":
    main()
