"""networks.models -> B200 parameter trees (same names as networks/models/__init__.py:1-11)."""
from aot_benchmark_b200.model import AOT, DeAOT, build_vos_model  # noqa: F401
