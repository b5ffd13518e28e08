import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
GOLDEN = os.path.join(TESTS, "golden")
for _p in (TESTS, GOLDEN):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """-m gpu tests must not silently pass without a device: fail loudly instead of skipping
    when they were explicitly selected; skip only in an unfiltered run on a CPU box."""
    import torch
    if torch.cuda.is_available():
        return
    selected = config.getoption("-m") or ""
    if "gpu" in selected and "not gpu" not in selected:
        return  # explicitly asked for: let them run and fail
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pixel_oracle():
    from oracle.dcnv3_oracle import PixelOracle
    return PixelOracle()


def _reload_knobs():
    """The library caches its DCNV3_B200_* environment knobs on first use (no getenv on the hot path);
    tests that switch kernel families in-process make it read them again."""
    try:
        from yolo_dual_b200 import _lib
        _lib.reload_knobs()
    except Exception:
        pass


@pytest.fixture(autouse=True)
def _fresh_knobs():
    # runs before every test, after the previous test's monkeypatch was undone: back to the real environment
    _reload_knobs()
    yield


@pytest.fixture
def monkeypatch(monkeypatch):
    """pytest's monkeypatch, with setenv / delenv of a DCNV3_B200_* name followed by a knob reload."""
    set_, del_ = monkeypatch.setenv, monkeypatch.delenv

    def setenv(name, value, *a, **k):
        set_(name, value, *a, **k)
        if name.startswith("DCNV3_B200_"):
            _reload_knobs()

    def delenv(name, *a, **k):
        del_(name, *a, **k)
        if name.startswith("DCNV3_B200_"):
            _reload_knobs()

    monkeypatch.setenv, monkeypatch.delenv = setenv, delenv
    return monkeypatch
