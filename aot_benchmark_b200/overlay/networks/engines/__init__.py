"""networks.engines -> B200 engines (same names as networks/engines/__init__.py:1-21)."""
from aot_benchmark_b200.engine import (AOTEngine, AOTInferEngine, DeAOTEngine,  # noqa: F401
                                       DeAOTInferEngine, build_engine)
