import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import rtw_pkg  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    m = rtw_pkg.load()
    from oracle.binding import ORACLE_LIB_PATH
    need = [m.api.RTW_LIB_PATH, ORACLE_LIB_PATH]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    return m


@pytest.fixture(scope="session")
def orc(pkg):
    """CPU oracle (checker)."""
    from oracle.binding import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def rtw(pkg):
    """CUDA product library (loads without a GPU; compute calls need one)."""
    return pkg.load_rtw()


@pytest.fixture(scope="session")
def gpu(rtw):
    if rtw.f("device_count")() < 1:
        pytest.fail("this test is marked gpu but no CUDA device is visible — librtw has no CPU fallback")
    return rtw


def q24(rs, shape):
    """U[0,1) draws on the 24-bit grid the device RNG uses (identical in f32 and f64)."""
    return rs.randint(0, 1 << 24, shape).astype(np.float64) / float(1 << 24)


def f32(a):
    """Round test inputs to f32-representable values so that oracle and device see identical numbers."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def record(key, **values):
    """Append a measured parity figure to gpurun_out/parity_measured.jsonl (comes back from the GPU box with gpurun):
    the bars in test_gpu_parity.py are set 1 % under what this file showed on the B200."""
    import json
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"key": key, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass
