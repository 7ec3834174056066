"""beatrice_vst_b200 -- B200-native drop-in for the beatricelib C API hot path.

The product is the C-ABI shared library ``csrc/libbeatrice_b200.so`` (CUDA,
sm_100a).  This Python package only holds tooling around it:

* ``model_spec``  -- the builder-defined network spec "M0"  and
  the seeded weight generator that writes a model directory in the layout
  ``ProcessorCore2::LoadModel`` expects (reference
  ``src/common/processor_core_2.cc:301-351``).
* ``lib``         -- ctypes bindings for the C-ABI (tests and ``bench.py``).
* ``signals``     -- the synthetic 48 kHz test signals of SURVEY.md section 8(d).

Nothing here computes audio on the CPU; if the CUDA library is missing the
bindings raise.
"""

__all__ = ["model_spec", "lib", "signals"]
