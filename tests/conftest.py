import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from falcon_unzip_b200 import engine
    return engine.get_engine(0)


_synth_cache = {}


def synth_set(name, **over):
    """Cached synthetic sets (generation is deterministic in the config)."""
    import dataclasses
    from falcon_unzip_b200 import synth
    key = (name, tuple(sorted(over.items())))
    if key not in _synth_cache:
        cfg = dataclasses.replace(synth.CONFIGS[name], **over)
        _synth_cache[key] = synth.generate(cfg)
    return _synth_cache[key]
