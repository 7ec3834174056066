"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU.

Nothing here (or in any ``-m gpu`` test) reads /root/reference at run time: the reference
call site is reached only through the prebuilt ``oracle/_ref/callsite_runner_*`` binaries.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec  # noqa: E402
import loader as oracle_loader  # noqa: E402  (oracle/loader.py: test infrastructure)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    import __graft_entry__ as g
    if not (os.path.exists(blib.PRODUCT_SO) and os.path.exists(oracle_loader.ORACLE_SO)):
        g.build()


@pytest.fixture(scope="session")
def model_dirs(tmp_path_factory):
    """{family: directory} of seeded synthetic models (seed 0, 8 speakers)."""
    out = {}
    for fam in (0, 1, 2):
        d = tmp_path_factory.mktemp(f"model_f{fam}")
        model_spec.write_model_dir(str(d), n_speakers=8, family=fam, seed=0)
        out[fam] = str(d)
    return out


@pytest.fixture(scope="session")
def model_dir(model_dirs):
    return model_dirs[2]


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    return oracle_loader.load_oracle()


@pytest.fixture(scope="session")
def product():
    _ensure_built()
    return blib.load_product()


def rms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))
