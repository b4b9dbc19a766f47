import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    from qwen3_tts_rs_b200 import lib as L
    return L.load()


_W = {}


def talker_weights(spec):
    from qwen3_tts_rs_b200 import weights as W
    if spec.name not in _W:
        _W[spec.name] = W.make_talker_weights(spec)
    return _W[spec.name]


_VW = {}


def vocoder_weights(vspec, key):
    from qwen3_tts_rs_b200 import weights as W
    if key not in _VW:
        _VW[key] = W.make_vocoder_weights(vspec)
    return _VW[key]
