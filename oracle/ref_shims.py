"""TEST INFRASTRUCTURE ONLY — import shims that let the *unmodified* reference
(`/root/reference`, yeyupiaoling/MASR @ fe0010de) import and run in this container.

Nothing here touches arithmetic.  The shims are the three kinds listed in
SURVEY.md §8(c):

1. empty ``sys.modules`` stubs for optional third-party packages that are only
   needed at import time (masr/data_utils/audio.py:7-8, masr/data_utils/utils.py:9-15,
   masr/utils/logger.py:5, masr/trainer.py:17, masr/utils/metrics.py:1);
2. ``torch.nn.modules.conv.{Union,Optional}`` re-exports
   (masr/model_utils/squeezeformer/conv2d.py:2 imports them from there);
3. ``np.sctypes`` (masr/data_utils/audio.py:542,567 use it; removed in NumPy 2).

The reference is only available in the build container (never on the GPU box), so
only ``tests/golden/make_golden.py`` and tests that are skipped when
``/root/reference`` is absent may call :func:`install`.
"""
import os
import sys
import types
import typing

REFERENCE_ROOT = os.environ.get("MASR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "masr"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        __import__(name)
        return sys.modules[name]
    except Exception:
        pass
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install():
    """Make ``import masr`` resolve to the read-only reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np
    import torch

    _stub("termcolor", colored=lambda s, *a, **k: s)
    _stub("resampy")
    _stub("soundfile")
    _stub("av")
    _stub("pydub", AudioSegment=object)
    _stub("zhconv", convert=lambda s, *a, **k: s)
    _stub("visualdl", LogWriter=object)
    _stub("Levenshtein")
    conv_mod = torch.nn.modules.conv
    if not hasattr(conv_mod, "Union"):
        conv_mod.Union = typing.Union
    if not hasattr(conv_mod, "Optional"):
        conv_mod.Optional = typing.Optional
    if not hasattr(np, "sctypes"):
        np.sctypes = {
            "float": [np.float16, np.float32, np.float64],
            "int": [np.int8, np.int16, np.int32, np.int64],
            "uint": [np.uint8, np.uint16, np.uint32, np.uint64],
            "complex": [np.complex64, np.complex128],
            "others": [bool, object, bytes, str, np.void],
        }
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import masr  # noqa: F401  (the reference package)
    return masr


def build_real_predictor(tmp: str, streaming: bool = True, wseed: int = 0, vocab_size: int = 4233):
    """The unmodified reference's ``masr.predict.MASRPredictor`` (conformer.yml, ctc_greedy, CPU) over the synthetic weights
    of ``masr_b200.synth`` — the same construction ``tests/golden/make_golden.py`` freezes its goldens from.  Used by
    ``bench.py --impl reference`` when a reference tree is present (build container; never on the GPU box)."""
    install()
    import numpy as np
    import torch
    import yaml
    from masr.model_utils.conformer.model import ConformerModel
    from masr.predict import MASRPredictor
    from masr_b200 import synth
    cfg = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "configs", "conformer.yml"), encoding="utf-8"))
    mi = os.path.join(tmp, f"mean_istd_{wseed}.json")
    synth.write_mean_istd(mi, wseed)
    model = ConformerModel(input_dim=80, vocab_size=vocab_size, mean_istd_path=mi, streaming=streaming,
                           encoder_conf=cfg["encoder_conf"], decoder_conf=cfg["decoder_conf"], **cfg["model_conf"])
    res = model.load_state_dict(synth.to_torch(synth.conformer_state_dict(wseed, vocab_size)), strict=False)
    assert not res.unexpected_keys
    mp = os.path.join(tmp, "inference.pt")
    torch.jit.save(model.eval().export(), mp)
    vp = os.path.join(tmp, "vocabulary.txt")
    synth.write_vocabulary(vp, vocab_size)
    cfg["dataset_conf"]["dataset_vocab"] = vp
    cfg["dataset_conf"]["mean_istd_path"] = mi
    cfg["decoder"] = "ctc_greedy"
    np.random.seed(0)
    return MASRPredictor(configs=cfg, model_path=mp, use_gpu=False)
