"""ORACLE LOADER -- test infrastructure only.

Loads ``oracle/libbeatrice_oracle.so`` (the CPU restatement of spec M0 behind the beatrice.h ABI) through the
same ctypes binding class the product uses.  Only ``tests/``, ``__graft_entry__.smoke()`` / ``build()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; nothing under
``beatrice_vst_b200/`` does.
"""
from __future__ import annotations

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from beatrice_vst_b200.lib import BeatriceLib  # noqa: E402

ORACLE_SO = os.path.join(_HERE, "libbeatrice_oracle.so")


def load_oracle() -> BeatriceLib:
    return BeatriceLib(ORACLE_SO)
