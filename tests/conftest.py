import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes view of libilqg_b200.so). Fails loudly when the library is not built."""
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle bindings (the checker)."""
    o = entry.load_oracle()
    o.lib()
    return o


MODEL_NAMES = ("inverted_pendulum", "hopper", "humanoid")


@pytest.fixture(scope="session")
def omodels(pkg, oracle):
    return {n: oracle.Model(os.path.join(pkg.MODELS_DIR, n + ".ilqgm")) for n in MODEL_NAMES}


@pytest.fixture(scope="session")
def handles(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    hs = {n: pkg.Handle(pkg.Model.named(n), 0) for n in ("inverted_pendulum", "hopper")}
    yield hs
    for h in hs.values():
        h.close()


def scenario_states(name, n, seed, oracle=None, om=None, roll=0):
    """Seeded knot inputs for a model; optionally rolled forward with the oracle so that warm starts and
    contacts are consistent (the scenario of /root/reference/tst/test_derivatives.cpp:38-47)."""
    rng = np.random.default_rng(seed)
    if name == "inverted_pendulum":
        q = rng.uniform(-0.5, 0.5, (n, 2)); v = rng.normal(0, 0.5, (n, 2)); u = rng.uniform(-1, 1, (n, 1))
    elif name == "hopper":
        q = np.zeros((n, 6)); q[:, 0] = rng.uniform(-.1, .1, n); q[:, 1] = rng.uniform(1.0, 1.4, n); q[:, 2] = rng.uniform(-.1, .1, n)
        q[:, 3:5] = rng.uniform(-1.9, -0.6, (n, 2)) * 0.5; q[:, 5] = rng.uniform(-.4, .4, n)
        v = rng.normal(0, 0.5, (n, 6)); u = rng.uniform(-1, 1, (n, 3))
    else:
        q = np.zeros((n, 28)); q[:, 2] = rng.uniform(1.2, 1.5, n); q[:, 3] = 1
        q[:, 3:7] += rng.normal(0, 0.05, (n, 4)); q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
        q[:, 7:] = rng.uniform(-.15, .15, (n, 21)); q[:, [13, 19]] -= 0.3
        v = rng.normal(0, 0.3, (n, 27)); u = rng.uniform(-.4, .4, (n, 21))
    w = np.zeros_like(v)
    if roll and oracle is not None:
        q, v, w, _ = oracle.step_batch(om, q, v, u, w, roll)
    return q, v, u, w
