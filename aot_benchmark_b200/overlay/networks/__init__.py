"""Drop-in overlay of the reference's ``networks`` package.

Put ``<repo>/aot_benchmark_b200/overlay`` *before* the reference checkout on PYTHONPATH:

    PYTHONPATH=<repo>:<repo>/aot_benchmark_b200/overlay:<reference> python <reference>/tools/eval.py ...

``networks.engines`` and ``networks.models`` then resolve to the B200 implementations below while
``networks.managers``, ``networks.layers``, ``dataloaders``, ``utils`` and ``configs`` keep coming
from the unedited reference: this package extends its search path with the
reference's ``networks`` directory found on sys.path.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(os.path.abspath(_p or "."), "networks")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        if os.path.isdir(os.path.join(_cand, "managers")):
            __path__.append(_cand)
