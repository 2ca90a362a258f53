"""valor_b200 — B200-native implementation of VALOR's tri-modal pretraining hot path.

Host side mirrors the reference's Python interface (model/pretrain.py: VALOR and the encoder
classes); all arithmetic runs in hand-written sm_100a kernels behind the C ABI declared in
include/valor_b200.h (csrc/libvalor_b200.so).  There is no CPU or library fallback.
"""
__all__ = ["VALOR", "default_opts"]


def __getattr__(name):
    if name in ("VALOR", "default_opts"):
        from . import pretrain
        return getattr(pretrain, name)
    raise AttributeError(name)
