"""aot_benchmark_b200 -- B200-native AOT/DeAOT mask-propagation hot path.

Public surface (mirrors the reference's seam):
    build_vos_model(name, cfg)            networks/models/__init__.py:5-11
    build_engine(name, phase, **kw)       networks/engines/__init__.py:5-21
    EngineConfig(exp, model)              configs/default.py:5-9
"""
from .configs import EngineConfig  # noqa: F401
from .model import build_vos_model  # noqa: F401


def build_engine(name, phase="train", **kwargs):
    from .engine import build_engine as _b
    return _b(name, phase=phase, **kwargs)
