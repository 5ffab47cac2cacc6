import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from pyhmmer_b200 import _lib
        _lib.context(0)
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    from pyhmmer_b200 import _lib
    return _lib.context(0)


@pytest.fixture(scope="session")
def amino():
    from pyhmmer_b200 import easel
    return easel.Alphabet.amino()


class ModelPair:
    """The same model on both sides: our OptimizedProfile and the reference configured from the same ASCII file."""

    def __init__(self, hmm, L=400):
        from pyhmmer_b200 import plan7
        from oracle import refshim
        self._tmp = tempfile.NamedTemporaryFile(suffix=".hmm")
        hmm.write(self._tmp)
        self._tmp.flush()
        self.path = self._tmp.name
        with plan7.HMMFile(self.path) as f:
            self.hmm = f.read()
        self.bg = plan7.Background(self.hmm.alphabet)
        self.profile = plan7.Profile(self.hmm.M, self.hmm.alphabet).configure(self.hmm, self.bg, L)
        self.om = self.profile.to_optimized()
        self.ref = refshim.RefModel(self.path, 0, L)


@pytest.fixture(scope="session")
def make_pair():
    return ModelPair
